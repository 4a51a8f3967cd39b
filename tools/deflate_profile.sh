python tools/deflate_timing.py 32 5
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/deflate_launches.csv python tools/deflate_timing.py 32 1 > gpurun_out/deflate_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/deflate_launches.csv")) if len(r)>10]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value"); u=h.index("Metric Unit")
for r in rows[1:]:
    if "deflate" in r[k]: print(r[k][:60], r[v], r[u])
PY
