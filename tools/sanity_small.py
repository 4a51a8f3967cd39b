"""small filter_streaks / process_img calls for compute-sanitizer (covers the TMA loaders, pad fills and the histogram)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
from pystripe import core, stack_stats
from tools import synth
img = synth.plane(9, (150, 180))
for mode, wav in (("wrap", "db5"), ("mean", "db5"), ("reflect", "db10"), ("linear_ramp", "db2")):
    out = core.filter_streaks(np.stack([img, img]), sigma=(20, 20), wavelet=wav, padding_mode=mode)
    print(mode, wav, int(out.sum()), flush=True)
out = core.filter_streaks(img, sigma=(12, 12), wavelet="coif8", padding_mode="reflect", bidirectional=True)
print("coif8", int(out.sum()), flush=True)
h = stack_stats.histogram(np.stack([img, img]))
assert h.sum() == 2 * img.size and np.array_equal(h, 2 * np.bincount(img.ravel(), minlength=65536))
print("hist ok")
