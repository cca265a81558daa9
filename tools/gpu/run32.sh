#!/bin/bash
python tools/block_clients_probe.py 10000 2>&1 | tail -4
