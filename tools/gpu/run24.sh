#!/bin/bash
# small tables: window-width sweep (FAB_WS = 5 / 6 / 7 / 8: rate, table size, build time) + registration after the scratch pre-allocation
for w in 5 6 7 8; do
  lib=fabric-mod_b200/lib/libfabgpu_ws$w.so; [ $w = 6 ] && lib=fabric-mod_b200/lib/libfabgpu_ecdsa.so
  echo "== FAB_WS=$w"; FABGPU_TRACE=1 SMALL_LIB=$lib python tools/small_bench.py 65536 262144 2>&1 | grep -v "^$"
done | tee gpurun_out/small_ws_sweep.txt
timeout 900 python -m pytest tests/test_gpu_small_tables.py tests/test_gpu_provider.py -m gpu -x -q 2>&1 | tail -4
