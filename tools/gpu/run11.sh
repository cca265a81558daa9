#!/bin/bash
# round-2 evidence on one GPU: bench line, launch list of the same command, full ncu captures of the dominant kernel and of SHA-256,
# and the whole GPU test suite
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
tail -c 600 gpurun_out/bench_r2_n1.err
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ecdsa_verify_cached_kernel -s 6 -c 1 -o gpurun_out/r2_final_cached python bench.py --steps 3 --warmup 3 --no-parity --no-block > gpurun_out/ncu_r2_cached.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sha256_segments_kernel -s 2 -c 3 -o gpurun_out/r2_sha256 python bench.py --steps 1 --warmup 3 --no-parity > gpurun_out/ncu_r2_sha.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2.txt 2>&1
tail -4 gpurun_out/pytest_gpu_r2.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','value_generic')}, d['e2e']['value'], d['e2e'].get('pageable_value'), d['e2e']['sync_value'], d.get('parity'))
PY
