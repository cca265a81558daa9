#!/bin/bash
# round-2 probe: Go toolchain on the GPU box, Jacobian-chain vs batch-affine key-table kernel, GPU test suite
(go version || echo "go: not found"; which go gccgo tinygo; ls /usr/local/go 2>&1) > gpurun_out/go_version.txt 2>&1
nvidia-smi -L >> gpurun_out/go_version.txt
FABGPU_CACHED_KERNEL=jac python tools/kbench.py fabric-mod_b200/lib/libfabgpu_ecdsa.so 65536 262144 > gpurun_out/kb_jac.txt 2>&1
FABGPU_CACHED_KERNEL=ba python tools/kbench.py fabric-mod_b200/lib/libfabgpu_ecdsa.so 65536 262144 > gpurun_out/kb_ba.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/kb_jac.txt gpurun_out/kb_ba.txt gpurun_out/pytest_gpu.txt
