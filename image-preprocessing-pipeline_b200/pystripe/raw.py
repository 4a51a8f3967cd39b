""".raw tiles: 8-byte header (width, height as uint32, either endianness) followed by uint16 pixels
(reference pystripe/raw.py:9-68)."""
import numpy as np


def raw_imread(path, dtype=None, shape=None):
    """memory-map a .raw tile; endianness is guessed from the header when dtype/shape are not given
    (the smaller of the two width interpretations wins, reference raw.py:20-38)."""
    if dtype is None or shape is None:
        head = np.fromfile(path, dtype=np.uint8, count=8)
        w_be, h_be = head.view(">u4")
        w_le, h_le = head.view("<u4")
        if w_le < w_be:
            shape, dtype = (int(h_le), int(w_le)), "<u2"
        else:
            shape, dtype = (int(h_be), int(w_be)), ">u2"
    return np.memmap(path, dtype=dtype, mode="r", offset=8, shape=tuple(shape))


def raw_imsave(path, img):
    """write header (width, height, native uint32) + uint16 pixels (reference raw.py:44-68)."""
    img = np.asarray(img)
    with open(path, "wb") as f:
        np.array([img.shape[1], img.shape[0]], dtype=np.uint32).tofile(f)
        img.astype(np.uint16, copy=False).tofile(f)
