#!/bin/bash
# bench with rotating inputs larger than the L2 (no fill between steps) + the tests that exercise the bucketed use counters
timeout 900 python -m pytest tests/test_gpu_small_tables.py tests/test_gpu_provider.py -m gpu -x -q 2>&1 | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1g.json 2> gpurun_out/bench_r2_n1g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1g.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','value_generic','value_small')}, d['config']['value_l2_fill_between_steps'], d['roofline_int']['frac'])
print({k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')}, d['clocks'])
PY
tail -2 gpurun_out/bench_r2_n1g.err
