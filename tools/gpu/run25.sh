#!/bin/bash
# FAB_WS = 8 default (segmented build): full GPU suite, smoke, registration trace, bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -1
FABGPU_TRACE=1 python tools/small_bench.py 65536 262144 2>&1 | tee gpurun_out/small_ws8.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1e.json 2> gpurun_out/bench_r2_n1e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1e.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','value_generic','value_small')}, d.get('small'))
print({k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')})
print(d['block_replay']['ms_per_block'], d['block_replay']['single_call']['ms_per_block'], d['parity']['mask_equals_oracle'])
PY
tail -3 gpurun_out/bench_r2_n1e.err
