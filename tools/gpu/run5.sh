#!/bin/bash
# prefetch variants, L2 fetch granularity, and an ncu capture of the batch-affine kernel at 256k
export KBENCH_ONLY=cached FABGPU_CACHED_KERNEL=ba
for g in 32 64 128; do
  echo "FABGPU_L2_FETCH=$g"; FABGPU_L2_FETCH=$g python tools/kbench.py fabric-mod_b200/lib/variants/cc_pf_k1.so 65536 262144 2>&1 | grep cached
done > gpurun_out/kb_ba_sweep3.txt
for f in fabric-mod_b200/lib/variants/cc_pf_k4.so fabric-mod_b200/lib/variants/cc_pf_mb3.so fabric-mod_b200/lib/variants/cc_pf_mb4.so; do
  python tools/kbench.py $f 65536 262144 2>&1 | grep cached
done >> gpurun_out/kb_ba_sweep3.txt
echo "jac with FABGPU_L2_FETCH=64 / 128" >> gpurun_out/kb_ba_sweep3.txt
FABGPU_CACHED_KERNEL=jac FABGPU_L2_FETCH=64 python tools/kbench.py fabric-mod_b200/lib/variants/cc_pf_k1.so 65536 262144 2>&1 | grep cached >> gpurun_out/kb_ba_sweep3.txt
FABGPU_CACHED_KERNEL=jac FABGPU_L2_FETCH=128 python tools/kbench.py fabric-mod_b200/lib/variants/cc_pf_k1.so 65536 262144 2>&1 | grep cached >> gpurun_out/kb_ba_sweep3.txt
cat gpurun_out/kb_ba_sweep3.txt
ncu --set full --clock-control none --import-source on -k regex:ecdsa_verify_ba_kernel -s 4 -c 1 -o gpurun_out/r2_ba_v3_256k python tools/kbench.py fabric-mod_b200/lib/variants/cc_pf_k1.so 262144 > gpurun_out/ncu_ba_v3.log 2>&1
tail -3 gpurun_out/ncu_ba_v3.log
