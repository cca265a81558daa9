#!/bin/bash
# sweep of the batch-affine kernel's CTA width (T) and inverter warps per round (K)
export KBENCH_ONLY=cached FABGPU_CACHED_KERNEL=ba
for f in fabric-mod_b200/lib/variants/ba_t*.so; do
  python tools/kbench.py $f 65536 262144 2>&1 | grep cached
done > gpurun_out/kb_ba_sweep1.txt
cat gpurun_out/kb_ba_sweep1.txt
