#!/bin/bash
# ncu --set full (+ source) of chosen kernels of one batch of a bench config:
#   tools/ncu_full.sh <config> <planes> name:regex:skip:count [...]
# raw / source pages come back as CSV under gpurun_out/ncu/; reports above 12 MB stay on the box.
set -u
cfg=$1; planes=$2; shift 2
mkdir -p gpurun_out/ncu
CMD="python tools/prof_workload.py $planes 1 $cfg"
$CMD > gpurun_out/ncu/plain_cfg$cfg.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu/plain_cfg$cfg.log; exit 1; }
for spec in "$@"; do
  IFS=: read name regex skip count <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -f -o gpurun_out/ncu/$name $CMD > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu/$name.source.csv 2>/dev/null
  gzip -f gpurun_out/ncu/$name.source.csv
  ls -la gpurun_out/ncu/$name.ncu-rep
  if [ $(stat -c %s gpurun_out/ncu/$name.ncu-rep) -gt 12000000 ]; then rm gpurun_out/ncu/$name.ncu-rep; fi
done
