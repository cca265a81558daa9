#!/bin/bash
# mixed batches with compacted generic signatures: GPU tests + the bench line (e2e.mixed_value_rank0)
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2b.txt 2>&1
tail -4 gpurun_out/pytest_gpu_r2b.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1b.json 2> gpurun_out/bench_r2_n1b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','value_generic')}, {k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')}, d['block_replay']['ms_per_block'])
PY
