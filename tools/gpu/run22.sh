#!/bin/bash
# small-table tier: launch shapes, registration cost, key-count (L2 residency) sweep
for th in 128 256 512; do FABGPU_SMALL_THREADS=$th python tools/small_bench.py 65536 262144 2>&1 | grep -v "^$" ; done | tee gpurun_out/small_shapes.txt
SMALL_KEYS=512 python tools/small_bench.py 65536 2>&1 | tee -a gpurun_out/small_shapes.txt
SMALL_KEYS=16000 python tools/small_bench.py 65536 2>&1 | tee -a gpurun_out/small_shapes.txt
