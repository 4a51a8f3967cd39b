"""process_img on a pageable numpy stack for several staging-thread counts (B2S_HOST_THREADS is read once per process)."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
    import time
    import numpy as np
    from pystripe import core
    from tools import synth
    base = synth.stack(8, (2048, 2048)); stack = np.concatenate([base] * 16)
    kw = dict(sigma=(256, 256), wavelet="db10", padding_mode="reflect", dark=100)
    for _ in range(3):
        r = core.process_img(stack, _max_batch=32, **kw)
    t = time.perf_counter()
    for _ in range(3):
        r = core.process_img(stack, _max_batch=32, **kw)
    dt = (time.perf_counter() - t) / 3
    print(f"B2S_HOST_THREADS={os.environ.get('B2S_HOST_THREADS')}: pageable {stack.shape[0] * 2048 * 2048 / dt / 1e6:8.0f} Mpx/s", flush=True)
else:
    for th in ("2", "4", "8", "12", "16"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, B2S_HOST_THREADS=th))
