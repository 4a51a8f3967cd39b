"""Golden vector for a whole stitched slice: the REFERENCE SOURCE VERBATIM (oracle/ref_runner.py) runs the
call `process_images.py:702-740` makes through `parallel_image_processor` — process_img with coif15, bidirectional,
sigma = 2 x tile side, lightsheet clean, 8-bit conversion, the bleach clip levels passed with frequency None — on a
seeded synthetic plane larger than a camera tile.  The full output is too large to commit: a strided sample (every 61st
row, every 67th column), four 64 x 64 corner/centre crops and a CRC32 of the whole array are stored.

    python tests/golden/make_golden_large.py [case ...]      (minutes of CPU; only where /root/reference exists)
"""
import json
import sys
import time
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_runner  # noqa: E402

if __name__ == "__main__":
    ref_runner.ensure_pinned_env()

import numpy as np  # noqa: E402

from tools import synth  # noqa: E402

CASES = {
    # name: (shape, kwargs of the stitched-slice call)
    "stitched_4096x6144_coif15_bidir_ls_8bit": ((4096, 6144), dict(
        sigma=(4096, 4096), wavelet="coif15", padding_mode="reflect", bidirectional=True, threshold=6.5,
        bleach_correction_frequency=None, bleach_correction_clip_min=4.9, bleach_correction_clip_med=6.5,
        bleach_correction_clip_max=8.1, bleach_correction_max_method=False, dark=100, lightsheet=True, percentile=0.25,
        rotate=0, convert_to_8bit=True, bit_shift_to_right=4, d_type="uint16")),
    "stitched_5000x3000_db9_wrap_16bit": ((5000, 3000), dict(
        sigma=(1024, 1024), wavelet="db9", padding_mode="wrap", bidirectional=True, dark=90, lightsheet=False,
        rotate=90, d_type="uint16")),
    "stitched_10000x14000_coif15_bidir_ls_8bit": ((10000, 14000), dict(
        sigma=(4096, 4096), wavelet="coif15", padding_mode="reflect", bidirectional=True, threshold=6.5,
        bleach_correction_frequency=None, bleach_correction_clip_min=4.9, bleach_correction_clip_med=6.5,
        bleach_correction_clip_max=8.1, dark=100, lightsheet=True, percentile=0.25, rotate=0,
        convert_to_8bit=True, bit_shift_to_right=3, d_type="uint16")),
}
BLOBS = {"stitched_10000x14000_coif15_bidir_ls_8bit": 24}      # plane synthesis costs n_blobs full-size passes


def digest(a: np.ndarray) -> dict:
    h, w = a.shape
    return {"sample": a[::61, ::67].copy(), "c00": a[:64, :64].copy(), "c01": a[:64, -64:].copy(),
            "c10": a[-64:, :64].copy(), "c11": a[-64:, -64:].copy(),
            "mid": a[h // 2 - 32:h // 2 + 32, w // 2 - 32:w // 2 + 32].copy()}


def plane_for(name):
    shape, kw = CASES[name]
    return synth.plane(7, shape, n_blobs=BLOBS.get(name, 120), seed=4321), dict(kw)


def main():
    core, _ = ref_runner.load()
    out, meta = {}, {}
    npz, js = ROOT / "tests" / "golden" / "large_plane_golden.npz", ROOT / "tests" / "golden" / "large_plane_golden.json"
    only = [a for a in sys.argv[1:] if a in CASES]
    if only and npz.exists():                     # update the named cases, keep the others
        with np.load(npz) as old:
            out = {k: old[k] for k in old.files}
        meta = json.loads(js.read_text())
    for name in only or CASES:
        img, kw = plane_for(name)
        kw["tile_size"] = img.shape
        t0 = time.perf_counter()
        res = core.process_img(img.copy(), **kw)
        dt = time.perf_counter() - t0
        for k, v in digest(res).items():
            out[f"{name}/{k}"] = v
        meta[name] = {"dtype": str(res.dtype), "shape": list(res.shape), "crc32": zlib.crc32(np.ascontiguousarray(res).tobytes()),
                      "reference_cpu_seconds": round(dt, 1)}
        print(name, meta[name], flush=True)
    np.savez_compressed(npz, **out)
    js.write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
