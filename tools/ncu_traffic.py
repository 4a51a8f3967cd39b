#!/usr/bin/env python3
"""profiles/r01_traffic.json from an `ncu --set full` capture of ONE batch of the bench workload
(`ncu --set full --clock-control none -k regex:'k_dwt_fwd|k_dwt_inv|k_notch_exact|k_prologue|k_epilogue' -o rep python
tools/prof_workload.py 32 1`): dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed the way bench.py names
kernels (class@Llevel).  Launch order of a batch: prologue, forward levels 1..L, notch levels 1..L, inverse levels L..1,
epilogue.

    python tools/ncu_traffic.py gpurun_out/main.ncu-rep 32 "profiles/r01_v10_main_kernels.md (...)" > profiles/r01_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

rep, planes, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}[u]


seen = {"dwt_fwd": [], "notch": [], "dwt_inv": [], "prologue": [], "epilogue": []}
for r in body:
    name = r[col["Kernel Name"]]
    cls = next((c for c in seen if c in name), None)
    if cls is None:
        continue
    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    seen[cls].append((b, to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])))
out = {"source": source, "kernels": {}}
levels = len(seen["dwt_fwd"])
for cls, lst in seen.items():
    for i, (b, us) in enumerate(lst):
        if cls in ("prologue", "epilogue"):
            key = cls
        else:
            lvl = levels - i if cls == "dwt_inv" else i + 1
            key = f"{cls}@L{lvl}"
        out["kernels"][key] = {"dram_bytes_per_plane": b / planes, "planes_per_launch": planes, "ncu_time_us_per_launch": us}
print(json.dumps(out, indent=1))
