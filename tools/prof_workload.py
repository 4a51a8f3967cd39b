"""One batch of a bench workload for ncu: tools/prof_workload.py <planes> <reps> [config]   (config = bench.py's --config, default 2:
process_img with flat + dark, 2048^2 u16, db10, sigma 256).  PW_SIGMA / PW_WAVELET / PW_NOFLAT override config 2's plan."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core, _native
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
config = int(sys.argv[3]) if len(sys.argv) > 3 else 2
base = synth.stack(4, (2048, 2048))
stack = torch.from_numpy(np.concatenate([base] * (n // 4))).cuda()
flat = core.normalize_flat(synth.flat_field((2048, 2048)))
out = None
if config == 2:
    sigma = tuple(int(v) for v in os.environ.get("PW_SIGMA", "256,256").split(","))      # PW_SIGMA=128,512 PW_WAVELET=coif15: config 5
    plan = core._get_plan(0, (2048, 2048), _native.U16, process=1, sigma=sigma, level=0, wavelet=os.environ.get("PW_WAVELET", "db10"),
                          threshold=None, padding_mode="reflect", bidirectional=False, log1p=True,
                          flat=None if os.environ.get("PW_NOFLAT") else flat, dark=100, out_code=_native.U16, max_batch=n)
    for _ in range(reps):
        out = plan.run_torch(stack, out)
else:
    import bench
    cfg = bench.CONFIGS[config]
    for _ in range(reps):
        if cfg["fn"] == "filter_streaks":
            out = core.filter_streaks(stack, **cfg["kw"])
        else:
            out = core.process_img(stack, flat=flat if cfg["flat"] else None, _max_batch=n, **cfg["kw"])
torch.cuda.synchronize()
print("ok", int(out[0, ::256, ::256].sum()))
