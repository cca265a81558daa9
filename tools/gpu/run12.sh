#!/bin/bash
# 8 GPUs: bench at N=8 (configs[3] parity leg: 1 048 576 signatures) and N=4 (configs[4]: 262 144), p2p exchange (NCCL timed beside it)
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-block > gpurun_out/bench_r2_n$n.json 2> gpurun_out/bench_r2_n$n.err
  tail -c 300 gpurun_out/bench_r2_n$n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_n$n.json').read().strip().splitlines()[-1])
print($n, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'nccl', d['config'].get('value_with_nccl_allgather'), 'e2e', d['e2e']['value'], d['e2e'].get('pageable_value'))
print(d.get('parity'))
PY
done
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
