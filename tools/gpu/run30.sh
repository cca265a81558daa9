#!/bin/bash
# 2 GPUs, final bench: NCCL all-gather as `value` (default), peer-memory exchange timed beside it; then the reverse
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --no-block > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2_final.json').read().strip().splitlines()[-1])
c=d['config']
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, c.get('collective'), c.get('value_with_peer_memory_exchange'), c.get('value_with_nccl_allgather'), c.get('value_l2_fill_between_steps'), c.get('collective_note'))
print({k:d['e2e'][k] for k in ('value','pageable_value','mixed_value_rank0')}, d['parity']['mask_equals_oracle'])
PY
tail -2 gpurun_out/bench_r2_n2_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 10 --warmup 3 --no-block --no-parity --collective p2p 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('p2p as value:', d['value'], c.get('collective'), c.get('value_with_nccl_allgather'))"
