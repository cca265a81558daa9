#!/bin/bash
# round-2 probe: Go toolchain on the GPU box, key-table kernel variants (Jacobian chain, lane-split x2 / x4, batch-affine), GPU tests
(go version || echo "go: not found"; which go gccgo tinygo; ls /usr/local/go 2>&1) > gpurun_out/go_version.txt 2>&1
nvidia-smi -L >> gpurun_out/go_version.txt
for v in jac l2 l4 ba; do
  FABGPU_CACHED_KERNEL=$v python tools/kbench.py fabric-mod_b200/lib/libfabgpu_ecdsa.so 65536 262144 > gpurun_out/kb_$v.txt 2>&1
done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
grep -h cached gpurun_out/kb_*.txt; tail -4 gpurun_out/pytest_gpu.txt
