"""Deterministic synthetic light-sheet planes (numpy only) — SURVEY.md §8(d).  Used by tests/ and bench.py."""
import numpy as np


def plane(z: int, shape=(2048, 2048), n_blobs: int = 40, seed: int = 1234) -> np.ndarray:
    """camera offset 110 + Gaussian blobs, multiplied by a per-row stripe gain, plus noise -> uint16."""
    h, w = shape
    rng = np.random.default_rng(seed + z)
    yy = np.arange(h, dtype=np.float32)[:, None]
    xx = np.arange(w, dtype=np.float32)[None, :]
    img = np.full(shape, 110.0, dtype=np.float32)
    scale = min(h, w) / 2048.0
    for _ in range(n_blobs):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        s = rng.uniform(4, 60) * max(scale, 0.05)
        a = rng.uniform(100, 4000)
        img += a * np.exp(-((yy - cy) ** 2) / (2 * s * s)) * np.exp(-((xx - cx) ** 2) / (2 * s * s))
    g = rng.standard_normal(h + 8).astype(np.float32)
    g = np.convolve(g, np.ones(9, dtype=np.float32) / 9, mode="valid")[:h]
    img *= (1 + 0.15 * g)[:, None]
    img += rng.normal(0, 5, size=shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16)


def stack(n: int, shape=(2048, 2048), seed: int = 1234, distinct: int = None) -> np.ndarray:
    """(n, H, W) uint16; `distinct` planes are generated and tiled to n (generation is ~0.5 s/plane at 2048^2)."""
    distinct = n if distinct is None else min(distinct, n)
    base = np.stack([plane(z, shape, seed=seed) for z in range(distinct)])
    if distinct == n:
        return base
    reps = -(-n // distinct)
    return np.concatenate([base] * reps)[:n]


def flat_field(shape=(2048, 2048)) -> np.ndarray:
    """0.6 + 0.4 exp(-r^2 / (2 (0.6 H)^2)), float32 (normalize_flat is applied by batch_filter)."""
    h, w = shape
    yy = (np.arange(h, dtype=np.float32) - (h - 1) / 2)[:, None]
    xx = (np.arange(w, dtype=np.float32) - (w - 1) / 2)[None, :]
    r2 = yy * yy + xx * xx
    return (0.6 + 0.4 * np.exp(-r2 / (2 * (0.6 * h) ** 2))).astype(np.float32)
