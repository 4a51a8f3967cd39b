"""pystripe.lightsheet_correct — module surface of the reference's pystripe/lightsheet_correct.py.

The windowed-percentile background clean runs inside process_img(lightsheet=True) on the GPU; this module keeps the
reference's public names.  `prctl` is the host helper process_images.py imports (process_images.py:39).
"""
import numpy as np


def prctl(data, percentiles):
    """lightsheet_correct.py:240-242 (numba np.percentile == numpy's default linear interpolation)."""
    return np.percentile(data, percentiles)


def correct_lightsheet(img, percentile=0.25, mask=None, lightsheet=dict(selem=(150, 1, 1)),
                       background=dict(selem=(200, 200, 1), spacing=(25, 25, 1), interpolate=1, dtype=None,
                                       step=(2, 2, 1)),
                       lightsheet_vs_background=2.0, return_lightsheet=False, return_background=False):
    """lightsheet_correct.py:31-106 for the parameterisation process_img uses (core.py:1333-1348):
    lightsheet selem (1, L, 1), background selem (B, B, 1), spacing (25, 25, 1), step (2, 2, 1)."""
    if mask is not None or return_lightsheet or return_background:
        raise NotImplementedError("mask / return_* options are outside the GPU hot path")
    from . import core
    ls_sel = tuple(lightsheet.get("selem", (150, 1, 1)))
    bg_sel = tuple(background.get("selem", (200, 200, 1)))
    if ls_sel[0] != 1 or bg_sel[0] != bg_sel[1] or tuple(background.get("spacing", (25, 25, 1)))[:2] != (25, 25) \
            or tuple(background.get("step", (2, 2, 1)))[:2] != (2, 2):
        raise NotImplementedError("only the structuring elements process_img passes are implemented on the GPU")
    d_type = lightsheet.get("dtype", None) or img.dtype
    return core.process_img(img, sigma=(0, 0), lightsheet=True, artifact_length=ls_sel[1],
                            background_window_size=bg_sel[0], percentile=percentile,
                            lightsheet_vs_background=lightsheet_vs_background, d_type=d_type)
