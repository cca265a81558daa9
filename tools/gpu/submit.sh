#!/bin/bash
# usage: tools/gpu/submit.sh <script under tools/gpu> <log name> [timeout seconds] [gpus]
# Submits one gpurun call, retrying while the pod answers "busy" (exit code 3, nothing charged).
script=$1; log=gpurun_out/$2.log; to=${3:-900}; gpus=${4:-1}
mkdir -p gpurun_out
for i in $(seq 1 20); do
  if [ "$gpus" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "bash tools/gpu/$script" > $log 2>&1; else /usr/local/graft/bin/gpurun --timeout $to -- "bash tools/gpu/$script" > $log 2>&1; fi
  rc=$?
  [ $rc -ne 3 ] && break
  sleep 90
done
echo "submit rc=$rc attempts=$i" >> $log
