"""Golden vector from the reference's ONLY real image: LsDeconvolveMultiGPU/supplements/test.png, a
2000 x 2000 uint16 light-sheet tile (values 0..432).  The REFERENCE SOURCE VERBATIM (oracle/ref_runner.py) runs the
Step-3 call of process_images.py:420-447 on it — process_img with db9, sigma = (250, 250), reflect padding, bidirectional,
uint16 output — plus the 8-bit variant the same script selects with need_16bit_to_8bit_conversion.

The input tile is data, not source: it is stored as a fixture (real_tile_input.npz, deflate-compressed) because the GPU
box has no /root/reference; outputs are stored as a strided sample + crops + CRC32 of the whole array.

    python tests/golden/make_golden_real_tile.py          (only where /root/reference exists)
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

GOLD = ROOT / "tests" / "golden"
SOURCE = Path("/root/reference/LsDeconvolveMultiGPU/supplements/test.png")
CASES = {
    # process_images.py:420-447 (tile_destriping_sigma from the GUI default 250)
    "real_tile_step3_db9_bidir_u16": dict(sigma=(250, 250), level=0, wavelet="db9", threshold=None,
                                          padding_mode="reflect", bidirectional=True, lightsheet=False, d_type="uint16",
                                          convert_to_8bit=False, bit_shift_to_right=8),
    "real_tile_step3_db9_bidir_8bit_shift2": dict(sigma=(250, 250), level=0, wavelet="db9", padding_mode="reflect",
                                                  bidirectional=True, d_type="uint16", convert_to_8bit=True,
                                                  bit_shift_to_right=2, dark=20),
}


def digest(a: np.ndarray) -> dict:
    h, w = a.shape
    return {"sample": a[::13, ::17].copy(), "c00": a[:64, :64].copy(), "c11": a[-64:, -64:].copy(),
            "mid": a[h // 2 - 64:h // 2 + 64, w // 2 - 64:w // 2 + 64].copy()}


def load_input() -> np.ndarray:
    with np.load(GOLD / "real_tile_input.npz") as z:
        return z["img"]


def main():
    from PIL import Image
    core, _ = ref_runner.load()
    with Image.open(SOURCE) as im:
        img = np.array(im)
    assert img.shape == (2000, 2000) and img.dtype == np.uint16
    np.savez_compressed(GOLD / "real_tile_input.npz", img=img)
    out, meta = {}, {"source": str(SOURCE.relative_to("/root/reference")), "input_crc32": zlib.crc32(img.tobytes())}
    for name, kw in CASES.items():
        t0 = time.perf_counter()
        res = core.process_img(img.copy(), tile_size=img.shape, **kw)
        dt = time.perf_counter() - t0
        for k, v in digest(res).items():
            out[f"{name}/{k}"] = v
        meta[name] = {"dtype": str(res.dtype), "shape": list(res.shape), "crc32": zlib.crc32(np.ascontiguousarray(res).tobytes()),
                      "reference_cpu_seconds": round(dt, 1), "changed_pixels_vs_input": int((res != img).sum()) if res.dtype == img.dtype else None}
        print(name, meta[name], flush=True)
    np.savez_compressed(GOLD / "real_tile_golden.npz", **out)
    (GOLD / "real_tile_golden.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
