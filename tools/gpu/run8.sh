#!/bin/bash
# L2 prefetch of the table entries (all leaves up front): ba, ba2 and the Jacobian-chain kernel
export KBENCH_ONLY=cached
{
for v in ba ba2; do
  for f in fabric-mod_b200/lib/variants/ee_bapf.so fabric-mod_b200/lib/variants/ee_bapf_pf.so; do
    echo "kernel=$v lib=$(basename $f)"; FABGPU_CACHED_KERNEL=$v python tools/kbench.py $f 65536 262144 1048576 2>&1 | grep -E "cached|rror"
  done
done
echo "kernel=jac lib=ee_jacpf.so (prefetch)"; FABGPU_CACHED_KERNEL=jac python tools/kbench.py fabric-mod_b200/lib/variants/ee_jacpf.so 65536 262144 1048576 2>&1 | grep -E "cached|rror"
echo "kernel=jac lib=ee_bapf.so (no prefetch)"; FABGPU_CACHED_KERNEL=jac python tools/kbench.py fabric-mod_b200/lib/variants/ee_bapf.so 65536 262144 1048576 2>&1 | grep -E "cached|rror"
} > gpurun_out/kb_l2prefetch.txt
cat gpurun_out/kb_l2prefetch.txt
