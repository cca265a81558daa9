#!/bin/bash
# two-signatures-per-thread batch-affine kernel (ba2): parity through the GPU tests that drive the key-table path, then timing
export FABGPU_CACHED_KERNEL=ba2
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config5 or key_tables or forced_key_tables or config2 or inplace or ragged" > gpurun_out/pytest_ba2.txt 2>&1
tail -3 gpurun_out/pytest_ba2.txt
export KBENCH_ONLY=cached
for f in fabric-mod_b200/lib/variants/dd_*.so; do
  python tools/kbench.py $f 65536 262144 1048576 2>&1 | grep -E "cached|rror"
done > gpurun_out/kb_ba2_sweep.txt
FABGPU_CACHED_KERNEL=ba python tools/kbench.py fabric-mod_b200/lib/variants/dd_nopf.so 65536 262144 1048576 2>&1 | grep -E "cached|rror" >> gpurun_out/kb_ba2_sweep.txt
FABGPU_CACHED_KERNEL=jac python tools/kbench.py fabric-mod_b200/lib/variants/dd_nopf.so 65536 262144 1048576 2>&1 | grep -E "cached|rror" >> gpurun_out/kb_ba2_sweep.txt
cat gpurun_out/kb_ba2_sweep.txt
