#!/usr/bin/env python3
"""Executed-instruction mix by SASS opcode from an ncu source-page CSV export.  python tools/ncu_opmix.py file.csv[.gz] [function-substr]"""
import collections, csv, gzip, io, sys
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else None
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
ops = collections.Counter(); fn = None; hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; iE = r.index("Instructions Executed"); continue
    if hdr is None or r[0] not in ("", "-") or len(r) <= iE: continue
    if want and want not in (fn or ""): continue
    sass = r[3].strip()
    if not sass: continue
    tok = sass.split()
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "LDGSTS")) and "." in op else "")
    try: ops[op] += int(r[iE] or 0)
    except ValueError: pass
tot = sum(ops.values())
print("total warp instructions", tot)
for op, n in ops.most_common(28):
    print(f"{op:14s} {n / tot:6.1%}")
