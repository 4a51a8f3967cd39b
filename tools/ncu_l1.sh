#!/bin/bash
# ncu --set full with source import of the level-1 kernels of one 32-plane batch of the bench workload.
# Reports stay on the box if they are large; the raw / source pages come back as CSV.
set -u
mkdir -p gpurun_out/ncu
CMD="python tools/prof_workload.py 32 1"
$CMD > gpurun_out/ncu/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu/plain.log; exit 1; }
prof() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c "$4" -f -o gpurun_out/ncu/$1 $CMD > gpurun_out/ncu/$1.log 2>&1
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu/$1.source.csv 2>/dev/null
  gzip -f gpurun_out/ncu/$1.source.csv
  ls -la gpurun_out/ncu/$1.ncu-rep
  if [ $(stat -c %s gpurun_out/ncu/$1.ncu-rep) -gt 12000000 ]; then rm gpurun_out/ncu/$1.ncu-rep; fi
}
for spec in "$@"; do
  IFS=: read name regex skip count <<< "$spec"
  prof "$name" "$regex" "$skip" "$count"
done
