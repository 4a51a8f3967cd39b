"""Whole stitched slices on one GPU: device time per plane of the post-stitch process_img call
(process_images.py:702-740) for the golden cases of tests/golden/make_golden_large.py, plus a 15000 x 20000 slice that
has no CPU golden (the reference needs ~20 min and > 60 GB for it).  Writes gpurun_out/large_report.json."""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core
from tests.golden import make_golden_large as gl
from tools import synth

cases = {n: gl.plane_for(n) for n in gl.CASES if n in sys.argv[1:] or len(sys.argv) == 1}
if len(sys.argv) == 1 or "full" in sys.argv[1:]:
    kw = dict(gl.CASES["stitched_10000x14000_coif15_bidir_ls_8bit"][1])
    cases["stitched_15000x20000_coif15_bidir_ls_8bit (no golden)"] = (synth.plane(3, (15000, 20000), n_blobs=8, seed=99), kw)
out = {}
for name, (img, kw) in cases.items():
    kw["tile_size"] = img.shape
    d_in = torch.from_numpy(img).cuda()[None]
    t0 = time.perf_counter()
    res = core.process_img(d_in, **kw)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        res = core.process_img(d_in, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    plan = next(iter(core._plans.values()))
    info = plan.info
    out[name] = {"shape": list(img.shape), "padded": [info.padded_height, info.padded_width], "levels": info.levels,
                 "ms_per_plane": round(ms, 2), "mpixel_per_s": round(img.size / ms / 1e3, 1),
                 "first_call_s": round(t_first, 2), "workspace_GB": round(info.workspace_bytes / 1e9, 2),
                 "out_dtype": str(res.dtype), "out_shape": list(res.shape[-2:])}
    print(name, json.dumps(out[name]), flush=True)
    del res, d_in
    core.clear_plan_cache()
    torch.cuda.empty_cache()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "large_report.json").write_text(json.dumps(out, indent=1))

# isotropic down-sampling of the same slices (parallel_image_processor.py:371-385), device-resident uint8 planes
from pystripe import isotropic as iso
iso_out = {}
for shape in ((4096, 6144), (10000, 14000), (15000, 20000)):
    t, m = iso.calculate_down_sampling_target(shape, shape, (1.0, 0.8, 0.8), 10.0)
    d_in = (torch.arange(shape[0] * shape[1], device="cuda", dtype=torch.int32) % 251).to(torch.uint8).reshape(1, *shape)
    iso.down_sample_xy(d_in, t, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        iso.down_sample_xy(d_in, t, m)
    e1.record()
    torch.cuda.synchronize()
    iso_out[f"{shape[0]}x{shape[1]}"] = {"target": list(t), "steps": len(m), "ms_per_plane": round(e0.elapsed_time(e1) / 3, 3)}
    print("isotropic", shape, iso_out[f"{shape[0]}x{shape[1]}"], flush=True)
    del d_in
out["isotropic_down_sample_xy"] = iso_out
(ROOT / "gpurun_out" / "large_report.json").write_text(json.dumps(out, indent=1))
