#!/bin/bash
# ncu --set full of the deflate kernels on 32 planes of 2048^2 (tools/deflate_timing.py)
set -u
mkdir -p gpurun_out/ncu
CMD="python tools/deflate_timing.py 32 1"
$CMD > gpurun_out/ncu/deflate_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/ncu/deflate_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_deflate' -s 8 -c 4 -f -o gpurun_out/ncu/deflate $CMD > gpurun_out/ncu/deflate.log 2>&1
ncu -i gpurun_out/ncu/deflate.ncu-rep --page raw --csv > gpurun_out/ncu/deflate.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/deflate.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu/deflate.source.csv 2>/dev/null
gzip -f gpurun_out/ncu/deflate.source.csv
ls -la gpurun_out/ncu/deflate.ncu-rep; rm -f gpurun_out/ncu/deflate.ncu-rep
