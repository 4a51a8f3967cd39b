"""Per-kernel-class CUDA-event times of one whole stitched slice (10000 x 14000 golden case) — where the 41 ms go."""
import sys, json
sys.path[:0] = ["/root/repo", "/root/repo/image-preprocessing-pipeline_b200"]
import torch
from pystripe import core, _native
from tests.golden import make_golden_large as gl
name = "stitched_10000x14000_coif15_bidir_ls_8bit"
img, kw = gl.plane_for(name); kw["tile_size"] = img.shape
d = torch.from_numpy(img).cuda()[None]
core.process_img(d, **kw); torch.cuda.synchronize()
ctx = _native.context(0)
ctx.timing_enable(True); ctx.timing_read(reset=True)
core.process_img(d, **kw)
tl = ctx.timing_read(reset=False, per_level=True)
t = ctx.timing_read(reset=True)
ctx.timing_enable(False)
print({k: round(v[0], 2) for k, v in t.items() if v[1]}, "ms")
print({f"{k}@L{l}": round(ms, 2) for (k, l), (ms, c) in sorted(tl.items()) if l and ms > 0.3})
