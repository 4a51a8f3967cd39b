"""Generate tests/golden/pystripe_golden.npz by running the REFERENCE SOURCE VERBATIM (oracle/ref_runner.py) on seeded
synthetic inputs.  Only runs where /root/reference exists (the build container); the .npz is committed.

    python tests/golden/make_golden.py

Inputs are not stored: tests regenerate them from the same seeds (tools/synth.py, numpy Generator).
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_runner  # noqa: E402

if __name__ == "__main__":
    ref_runner.ensure_pinned_env()

import numpy as np  # noqa: E402

from tests.golden import cases  # noqa: E402


def main():
    core, ls = ref_runner.load()
    out = {}
    meta = {}
    for name, kind, img, kw in cases.all_cases():
        if kind == "filter_streaks":
            res = core.filter_streaks(img.copy(), **kw)
        else:
            kw = dict(kw)
            flat = kw.pop("_flat", None)
            if flat is not None:
                # the reference only works with a flat when the caller passes float32 (core.py:1250 as written)
                res = core.process_img(img.astype(np.float32), flat=flat, d_type="uint16", **kw)
            else:
                res = core.process_img(img.copy(), **kw)
        out[name] = res
        meta[name] = {"kind": kind, "dtype": str(res.dtype), "shape": list(res.shape)}
    np.savez_compressed(ROOT / "tests" / "golden" / "pystripe_golden.npz", **out)
    (ROOT / "tests" / "golden" / "pystripe_golden.json").write_text(json.dumps(meta, indent=1))
    print(f"wrote {len(out)} golden outputs")


if __name__ == "__main__":
    main()
