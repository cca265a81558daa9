#!/bin/bash
# ncu of the batch-affine kernel at 64k (full set, one launch) + the GPU test suite
export FABGPU_CACHED_KERNEL=ba
ncu --set full --clock-control none --import-source on -k regex:ecdsa_verify_ba_kernel -s 4 -c 1 -o gpurun_out/r2_ba_v1 python tools/kbench.py fabric-mod_b200/lib/libfabgpu_ecdsa.so 65536 > gpurun_out/ncu_ba_v1.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1
tail -8 gpurun_out/pytest_gpu.txt
