#!/bin/bash
# bench line only (mixed-batch compaction with slot-owned scratch)
python bench.py --gpus 1 --steps 20 --warmup 5 --no-block --no-parity > gpurun_out/bench_r2_n1c.json 2> gpurun_out/bench_r2_n1c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','value_generic')}, {k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')})
PY
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mixed or eviction or forced or large_call" 2>&1 | tail -2
