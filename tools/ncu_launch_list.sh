#!/bin/bash
# ncu launch list (device time of every kernel launch) of a short bench.py run: tools/ncu_launch_list.sh <name> <bench args...>
set -u
name=$1; shift
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 1 --planes 32 --no-cpu-baseline --no-batch-filter $*"
$CMD > gpurun_out/ncu/$name.plain.json 2> gpurun_out/ncu/$name.plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu/$name.plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu/$name.launches.csv $CMD > gpurun_out/ncu/$name.ncu.log 2>&1
python - <<PY
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/ncu/$name.launches.csv")) if len(r) > 10]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value"); iU=hdr.index("Metric Unit")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[iV].replace(",",""))
    except ValueError: continue
    u=r[iU]
    v = v/1000 if u in ("ns","nsecond") else (v*1000 if u in ("ms","msecond") else v)   # -> us
    k=re.sub(r"<.*","",r[iK].replace("<unnamed>::","").replace("void ",""))
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share |"); print("|---|---|---|---|")
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"| {k} | {n} | {t:.0f} | {t/tot:.1%} |")
PY
gzip -f gpurun_out/ncu/$name.launches.csv
