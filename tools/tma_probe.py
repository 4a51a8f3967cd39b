import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
from pystripe import core
from tools import synth
img = synth.plane(6, (160, 200))
out = core.filter_streaks(img, sigma=(64, 64), wavelet="db2")
print("ok", out.shape, int(out.sum()))
