#!/bin/bash
# re-entry check of HEAD: full GPU suite, smoke, bench line, fe29 microbenchmark
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1d.json 2> gpurun_out/bench_r2_n1d.err
tail -c 3000 gpurun_out/bench_r2_n1d.json
cd profiles/microbench/fe29 && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../../fabric-mod_b200/csrc -o /tmp/bench29 bench29.cu && /tmp/bench29 | tee ../../../gpurun_out/fe29.txt
