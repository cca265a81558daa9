#!/bin/bash
# 2 GPUs: the tests that need more than one device, then the driver's N=2 command
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_parity.py -m gpu -q -k "peer or multi_device or one_million" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2_final.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','value_small')}, d['config'].get('value_with_nccl_allgather'), d['config'].get('collective_note'))
print({k:d['e2e'][k] for k in ('value','pageable_value','mixed_value_rank0')}, d['parity'])
PY
tail -3 gpurun_out/bench_r2_n2_final.err
