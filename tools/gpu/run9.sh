#!/bin/bash
./profiles/microbench/gather > gpurun_out/gather_b200.txt 2>&1; cat gpurun_out/gather_b200.txt
