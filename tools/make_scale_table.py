"""profiles/r02_scale/README.md from the bench lines kept under profiles/ (config 2 at N = 1 is profiles/r02_bench_final.json)."""
import json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
P = ROOT / "profiles"


def load(k, n):
    p = P / "r02_bench_final.json" if (k == 2 and n == 1) else P / "r02_scale" / f"config{k}_n{n}.json"
    return json.load(open(p)) if p.exists() else None


out = ["# r02 — BASELINE configs 2-5 at 1 … 8 B200s (`bench.py --config K --gpus N --steps 4 --warmup 3`; N > 1 under torchrun, one rank per GPU), final round-2 code", "",
       "All values Mpixel/s, whole job.  `value`: CUDA tensors in / out; `e2e`: page-locked host arrays through the public API; `pageable`: ordinary numpy stacks; `batch_filter`: TIFF files on tmpfs in and out — stored output, and the reference's default `('ADOBE_DEFLATE', 1)` output with the strips deflated on the GPU (`b2s_deflate_strips`).  Efficiency = N-GPU number / (N × the 1-GPU number of the same leg, another box).", "",
       "| config | N | value | e2e | pageable | batch_filter stored | batch_filter deflate-1 (GPU encoder) | value eff. | e2e eff. | CPU port Mpx/s (cores) |", "|---|---|---|---|---|---|---|---|---|---|"]
for k in (2, 3, 4, 5):
    base = None
    for n in (1, 2, 4, 8):
        d = load(k, n)
        if d is None:
            continue
        b = d["e2e_batch_filter"]
        v, e = d["value"], d["e2e"]["value"]
        if n == 1:
            base = (v, e)
        ve = f"{v / (n * base[0]):.2f}" if n > 1 else "—"
        ee = f"{e / (n * base[1]):.2f}" if n > 1 else "—"
        cpu = d.get("cpu_baseline") or {}
        cpus = f"{cpu['value']:.1f} ({cpu['cores']})" if cpu.get("value") else "—"
        out.append(f"| {k} | {n} | {v:,.0f} | {e:,.0f} | {d['e2e_public_api']['value']:,.0f} | {b['uncompressed']['value']:,.0f} | {b['adobe_deflate_1']['value']:,.0f} | {ve} | {ee} | {cpus} |")
out += ["", "Config 2 at N = 1 is `profiles/r02_bench_final.json` (the default `python bench.py`).  End to end at N = 8 is bounded by the 8-GPU box's host side (about 71 GB/s per direction in total, r01 `tools/pcie_probe_multi.py`; 32 host cores for 8 ranks): config 3 returns uint8 quarter-size planes (1/8 of the D2H bytes), config 5 is compute-heavy enough (coif15, two passes) that eight GPUs still scale at 0.92.  With the GPU deflate encoder the default-compression `batch_filter` leg went from 0.8 (host zlib, first half of round 2) to 13–23 Gpixel/s at N = 8 — it beats the stored leg because fewer bytes reach the files.  Config 2 at N = 2 and N = 4 ran on one 4-GPU box (the N = 4 run used all of it): its host side moves less than the 8-GPU box's, so e2e at N = 4 is below N = 2 — the e2e curve beyond one GPU measures the box, not the kernels.  The N = 1 lines and config 2 at N = 8 were re-measured after the staging copies moved to non-temporal stores: +33 % for pageable stacks on one GPU (eight staging threads), nothing at N = 8 (two staging threads per rank on a host that is saturated either way); the other N > 1 lines predate that change.  Box-to-box variance of the host-bound legs is about ±5 % (single-GPU e2e over the round's boxes: 20.5 – 22.8)."]
(P / "r02_scale" / "README.md").write_text("\n".join(out) + "\n")
print("\n".join(out[4:16]))
