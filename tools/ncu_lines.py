#!/usr/bin/env python3
"""Per-source-line instruction counts of the kernel in an .ncu-rep (needs -lineinfo, --set full --import-source on)."""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
per = collections.Counter(); samp = collections.Counter(); src = {}
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] not in ("", "-"):
        cur = int(r[0]); src[cur] = r[1]
        continue
    try: n = int(r[iE])
    except Exception: continue
    per[cur] += n; samp[cur] += int(r[iS] or 0)
tot = sum(per.values()); ts = sum(samp.values())
print("total", tot, "samples", ts)
for ln, n in per.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{ln:5d} {n / tot:6.1%} inst {samp[ln] / max(ts, 1):6.1%} samp | {src.get(ln, '')[:110]}")
