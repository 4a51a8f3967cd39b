#!/usr/bin/env python3
"""Opcode mix / hottest SASS instructions of the (single) kernel in an .ncu-rep (needs --set full --import-source on)."""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, body = rows[1], rows[2:]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW, iWI = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
ops, samp, tot, wf, wfi, ts = collections.Counter(), collections.Counter(), 0, 0, 0, 0
for r in body:
    try: n = int(r[iE])
    except Exception: continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += n; samp[op] += int(r[iSamp]); tot += n; ts += int(r[iSamp])
    wf += int(r[iW] or 0); wfi += int(r[iWI] or 0)
print("total warp inst", tot, "| smem wavefronts", wf, "ideal", wfi, "| samples", ts)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{op:10s} {n:12d} {n / tot:6.1%}  samples {samp[op] / max(ts, 1):6.1%}")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[hdr.index(h)] or 0) for r in body if len(r) > hdr.index(h) and r[hdr.index(h)].isdigit()) for h in stalls}
print("stalls:", ", ".join(f"{k[6:]}={v / max(ts, 1):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
