#!/bin/bash
# round-end measurement sequence on one B200 (gpurun): tests, bench (both arms), smoke, launch list, ncu --set full.
set -u
O=gpurun_out/final; mkdir -p $O gpurun_out/ncu
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
cp gpurun_out/parity_report*.json $O/ 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('bench', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['e2e_batch_filter']['adobe_deflate_1']['value'])"
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference.json; echo
tools/ncu_launch_list.sh final > $O/launch_list.md 2>&1; head -14 $O/launch_list.md
CMD="python tools/prof_workload.py 32 1"
ncu --set full --clock-control none -k regex:'k_dwt_fwd|k_dwt_inv|k_notch_exact|k_prologue|k_epilogue' -c 23 -f -o gpurun_out/ncu/final_main $CMD > gpurun_out/ncu/final_main.log 2>&1
ncu -i gpurun_out/ncu/final_main.ncu-rep --page raw --csv > $O/final_main.raw.csv 2>/dev/null
ls -la gpurun_out/ncu/final_main.ncu-rep; rm -f gpurun_out/ncu/final_main.ncu-rep
for k in 3 4 5; do python bench.py --config $k --steps 4 --warmup 3 > $O/config${k}_n1.json 2> $O/config${k}_n1.err; python -c "
import json; d=json.load(open('$O/config${k}_n1.json')); print('config $k', round(d['value']), round(d['e2e']['value']))"; done
