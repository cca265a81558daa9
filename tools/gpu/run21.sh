#!/bin/bash
# small-table tier: its GPU tests, the suites that share code with it, a bench line with value_small
timeout 900 python -m pytest tests/test_gpu_small_tables.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_block.py tests/test_gpu_provider.py -m gpu -x -q -k "not one_million" 2>&1 | tail -5
python bench.py --gpus 1 --steps 20 --warmup 5 --no-parity > gpurun_out/bench_r2_small1.json 2> gpurun_out/bench_r2_small1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_small1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','value_generic','value_small')}, d.get('small'))
print({k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')})
print(d['block_replay']['ms_per_block'], d['block_replay']['single_call'])
PY
tail -5 gpurun_out/bench_r2_small1.err
