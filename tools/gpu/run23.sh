#!/bin/bash
# small-table tier: where the registration time goes (FABGPU_TRACE), ncu capture of the kernel, launch list of the build kernels
FABGPU_TRACE=1 python tools/small_bench.py 65536 262144 2>&1 | tee gpurun_out/small_trace.txt
FABGPU_TRACE=1 SMALL_KEYS=16000 python tools/small_bench.py 65536 2>&1 | tee -a gpurun_out/small_trace.txt
ncu --set full --clock-control none --import-source on -k regex:ecdsa_verify_small_kernel -s 2 -c 1 -f -o gpurun_out/r2_small python tools/ncu_small_target.py > gpurun_out/ncu_small.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_small.ncu-rep > gpurun_out/r2_small_ncu_summary.txt 2>&1
head -40 gpurun_out/r2_small_ncu_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:small_ --csv --log-file gpurun_out/small_build_launches.csv python tools/ncu_small_target.py > /dev/null 2>&1
tail -8 gpurun_out/small_build_launches.csv
