"""pystripe — B200-native drop-in for the reference's pystripe package (same import surface as pystripe/__init__.py:1-2)."""
from .core import (filter_streaks, batch_filter, np_gaussian_filter, hist_match, max_level, foreground_fraction,  # noqa: F401
                   imread_tif_raw_png, imread_dcimg, imsave_tif, normalize_flat, process_img, process_stack,
                   read_filter_save)
