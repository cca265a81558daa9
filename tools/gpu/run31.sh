#!/bin/bash
# MSP with more identities than window-table slots (pre-registered peers keep their window tables) + the block-replay leg with 2 000 clients
timeout 600 python -m pytest tests/test_gpu_small_tables.py tests/test_gpu_block.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 --no-parity > gpurun_out/bench_r2_n1h.json 2> gpurun_out/bench_r2_n1h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1h.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_small')}, d['e2e']['value'])
b=d['block_replay']; print(b['ms_per_block'], b['single_call']['ms_per_block'], b.get('many_clients'))
PY
tail -3 gpurun_out/bench_r2_n1h.err
