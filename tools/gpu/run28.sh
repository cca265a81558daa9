#!/bin/bash
# final-state check on one GPU: full suite, smoke (now incl. the small-table kernel), the driver's bench command, launch list of the bench command
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_n1f.json 2> gpurun_out/bench_r2_n1f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1f.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','value_generic','value_small')})
print({k:d['e2e'][k] for k in ('value','pageable_value','sync_value','mixed_value_rank0')}, d['small']['tables'])
PY
tail -2 gpurun_out/bench_r2_n1f.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>/dev/null | tail -c 600
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-block --no-parity > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_final.csv')) if len(r)>14 and r[0].isdigit()]
t=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split('(')[0]; t[k][0]+=1; t[k][1]+=float(r[14])/1e3
for k,(n,us) in sorted(t.items(), key=lambda kv:-kv[1][1])[:14]: print("%-40s %5d launches %10.1f us total %8.1f us each" % (k,n,us,us/n))
PY
