#!/bin/bash
# 2 GPUs: peer-memory bitmask exchange test, multi-device context test, bench at N=2 (p2p and nccl collectives) with the parity leg
python -m pytest tests/test_gpu_peer.py tests/test_gpu_parity.py -m gpu -q -k "peer or multi_device" > gpurun_out/pytest_n2.txt 2>&1
tail -5 gpurun_out/pytest_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-block > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
tail -3 gpurun_out/bench_r2_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['parallelism'][:60], d['config'].get('value_with_nccl_allgather'), d['e2e']['value'], d['e2e'].get('pageable_value'), d.get('parity'))
PY
