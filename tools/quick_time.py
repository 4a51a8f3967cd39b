"""Developer timing loop (GPU box): per-kernel-class CUDA-event times for the headline plan."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core, _native
from tools import synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    wavelet = sys.argv[3] if len(sys.argv) > 3 else "db10"
    exact = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    base = synth.stack(4, (2048, 2048))
    stack = torch.from_numpy(np.concatenate([base] * (n // 4))).cuda()
    import os
    sigma = tuple(int(v) for v in os.environ.get("QT_SIGMA", "256,256").split(","))
    plan = core._get_plan(0, (2048, 2048), 1, process=0, sigma=sigma, level=0, wavelet=wavelet, threshold=None,
                          padding_mode=os.environ.get("QT_PAD", "wrap"), bidirectional=bool(int(os.environ.get("QT_BIDIR", "0"))),
                          log1p=True, max_batch=batch, exact=exact)
    ctx = plan.ctx
    out = plan.run_torch(stack); torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_read(reset=True)
    for _ in range(3):
        plan.run_torch(stack, out)
    tl = ctx.timing_read(reset=False, per_level=True)
    t = ctx.timing_read(reset=True)
    ctx.timing_enable(False)
    for (k, l), (ms, cnt) in sorted(tl.items()):
        if l: print(f"  {k}@L{l}: {ms / (3 * n) * 1e3:7.1f} us/plane")
    tot = sum(v[0] for v in t.values())
    for k, (ms, cnt) in t.items():
        if cnt: print(f"{k:10s} {ms / (3 * n) * 1e3:9.1f} us/plane  launches={cnt}")
    print(f"sum {tot / (3 * n) * 1e3:.1f} us/plane")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.run_torch(stack, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"end-to-end device-resident: {ms / n * 1e3:.1f} us/plane  {n * 2048 * 2048 / ms / 1e3:.1f} Mpx/s  "
          f"alg GB/s={plan.info.algorithmic_bytes_per_plane * n / ms / 1e6:.0f}")

main()
