#!/bin/bash
PROBE_FAST=1 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'block_|sha256|compact|verify' -c 60 --csv --log-file gpurun_out/block_clients_launches.csv python tools/block_clients_probe.py 10000 > gpurun_out/block_clients_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/block_clients_launches.csv')) if len(r)>14 and r[0].isdigit()]
for r in rows: print(r[0], r[4].split('(')[0][:40], r[7], r[8], round(float(r[14])/1e3,1))
PY
