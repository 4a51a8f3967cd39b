#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export (gzip ok): instructions executed,
stall samples and the dominant stall reasons per line.   python tools/ncu_src_summary.py file.csv[.gz] [top] [function-substr]"""
import collections, csv, gzip, io, sys
import os
SORTKEY=os.environ.get("SORTKEY","samp")
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
rows = list(csv.reader(io.StringIO(raw)))
sections, cur, fname, fpath = [], None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]; continue
    if r[0] == "Function Name":
        fname = r[1]; continue
    if r[0] == "Line No":
        cur = {"hdr": r, "rows": [], "file": fpath, "fn": fname}; sections.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
agg = collections.defaultdict(lambda: {"inst": 0, "samp": 0, "stalls": collections.Counter(), "src": "", "conf": 0})
seen_fn = set()
for s in sections:
    if want and want not in (s["fn"] or ""):
        continue
    seen_fn.add(s["fn"])
    h = s["hdr"]
    iE, iS = h.index("Instructions Executed"), h.index("# Samples")
    iC = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None
    st_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    line = None
    for r in s["rows"]:
        if r[0] not in ("", "-"):
            line = (s["file"].split("/")[-1], int(r[0])); agg[line]["src"] = r[1]; continue
        if line is None or len(r) <= iE:
            continue
        try:
            agg[line]["inst"] += int(r[iE] or 0); agg[line]["samp"] += int(r[iS] or 0)
            if iC is not None: agg[line]["conf"] += int(r[iC] or 0)
            for i, c in st_cols:
                v = int(r[i] or 0)
                if v: agg[line]["stalls"][c[6:]] += v
        except ValueError:
            pass
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print("functions:", sorted(x for x in seen_fn if x)[:4], "total inst", ti, "samples", ts)
allst = collections.Counter()
for a in agg.values(): allst.update(a["stalls"])
print("stall mix:", ", ".join(f"{k} {v / max(1, sum(allst.values())):.1%}" for k, v in allst.most_common(8)))
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in a["stalls"].most_common(3))
    print(f"{line[0][:14]:14s}:{line[1]:5d} inst {a['inst'] / max(ti, 1):6.1%} samp {a['samp'] / max(ts, 1):6.1%} conf {a['conf']:8d} | {a['src'][:70]:70s} | {st}")
