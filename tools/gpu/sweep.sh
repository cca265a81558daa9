#!/bin/bash
# usage: sweep.sh <glob under fabric-mod_b200/lib/variants> <out name> [batches...]: key-table kernel timing of every matching library variant
pat=$1; out=$2; shift 2; batches=${@:-65536 262144}
export KBENCH_ONLY=cached FABGPU_CACHED_KERNEL=${FABGPU_CACHED_KERNEL:-jac}
for f in fabric-mod_b200/lib/variants/$pat; do
  python tools/kbench.py $f $batches 2>&1 | grep -E "cached|Error|error" 
done > gpurun_out/$out.txt
cat gpurun_out/$out.txt
