#!/bin/bash
# 8 GPUs: the driver's N=8 command (block-replay leg skipped): peer-memory exchange after the one-fence-per-CTA fix vs the NCCL all-gather
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-block > gpurun_out/bench_r2_n8_final.json 2> gpurun_out/bench_r2_n8_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n8_final.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['config'].get('value_with_nccl_allgather'), d['config'].get('collective_note'))
print({k:d['e2e'][k] for k in ('value','pageable_value','mixed_value_rank0')}, {k:d['parity'][k] for k in ('n','zeros','mask_equals_oracle','status_equals_oracle')})
PY
tail -3 gpurun_out/bench_r2_n8_final.err
