#!/usr/bin/env python3
"""Generate orthogonal wavelet decomposition low-pass tables (dec_lo) from first principles.

PyWavelets is not installed anywhere in this environment, so the tables that
`pywt.Wavelet(name).dec_lo` would return are recomputed here with mpmath at 80 digits:

  dbN   : spectral factorisation of the Daubechies half-band polynomial, minimum-phase roots.
  symN  : same polynomial, root subset chosen for the least-asymmetric phase (Daubechies' criterion).
  coifN : Newton solve (mpmath.findroot) of the coiflet moment equations, 6N taps.

Convention (checked against the known-answer vectors in SURVEY.md §8c): the list written is pywt's
`dec_lo`; rec_lo = reverse(dec_lo); rec_hi[k] = (-1)^k rec_lo[F-1-k]; dec_hi = reverse(rec_hi).

Outputs (same numbers, two consumers):
  image-preprocessing-pipeline_b200/pystripe/_wavelet_tables.py   (product, host side)
  oracle/wavelet_tables.json                                      (oracle side)
"""
import json
import sys
from pathlib import Path

import mpmath as mp

mp.mp.dps = 80
ROOT = Path(__file__).resolve().parent.parent


def _halfband_roots(N):
    """Roots y_k of P(y) = sum_{k<N} C(N-1+k,k) y^k, mapped to z (both z_k and 1/z_k returned as pairs)."""
    if N == 1:
        return []
    coeffs = [mp.binomial(N - 1 + k, k) for k in range(N)]  # ascending
    ys = mp.polyroots(coeffs[::-1], maxsteps=2000, extraprec=2000)
    pairs = []
    for y in ys:
        b = 2 - 4 * y
        disc = mp.sqrt(b * b - 4)
        z1 = (b + disc) / 2
        z2 = (b - disc) / 2
        zin, zout = (z1, z2) if abs(z1) < 1 else (z2, z1)
        pairs.append((zin, zout))
    return pairs


def _poly_from_roots(N, roots):
    """ascending coefficients of (1+z)^N * prod (z - r), normalised to sum sqrt(2)."""
    poly = [mp.mpf(1)]
    for _ in range(N):  # times (1 + z)
        poly = [(poly[i] if i < len(poly) else 0) + (poly[i - 1] if i >= 1 else 0) for i in range(len(poly) + 1)]
    for r in roots:  # times (z - r)
        poly = [((poly[i - 1] if i >= 1 else 0) - r * (poly[i] if i < len(poly) else 0)) for i in range(len(poly) + 1)]
    poly = [mp.re(c) for c in poly]
    s = mp.fsum(poly)
    return [c * mp.sqrt(2) / s for c in poly]


def daubechies(N):
    roots = [zin for zin, _ in _halfband_roots(N)]
    return _poly_from_roots(N, roots)


def _group_conjugates(pairs):
    """group the (zin, zout) pairs into real singles and complex-conjugate couples."""
    used = [False] * len(pairs)
    groups = []
    for i, (zi, zo) in enumerate(pairs):
        if used[i]:
            continue
        used[i] = True
        if abs(mp.im(zi)) < mp.mpf(10) ** (-40):
            groups.append([(mp.re(zi), mp.re(zo))])
        else:
            j = min((k for k in range(len(pairs)) if not used[k]), key=lambda k: abs(pairs[k][0] - mp.conj(zi)))
            used[j] = True
            groups.append([(zi, zo), pairs[j]])
    return groups


def _phase_nonlinearity(h):
    """deviation of the unwrapped phase of H(w) from a straight line (Daubechies' least-asymmetric criterion)."""
    import numpy as np
    hh = np.array([float(c) for c in h])
    w = np.linspace(0, np.pi, 513)[1:-1]
    H = np.array([np.sum(hh * np.exp(-1j * ww * np.arange(len(hh)))) for ww in w])
    ph = np.unwrap(np.angle(H))
    A = np.vstack([w, np.ones_like(w)]).T
    res = ph - A @ np.linalg.lstsq(A, ph, rcond=None)[0]
    return float(np.sum(res ** 2))


def symlet(N):
    """least-asymmetric Daubechies filter: enumerate inside/outside choices per conjugate group."""
    pairs = _halfband_roots(N)
    groups = _group_conjugates(pairs)
    best = None
    for mask in range(1 << len(groups)):
        roots = []
        for g, grp in enumerate(groups):
            pick = (mask >> g) & 1
            roots.extend(p[pick] for p in grp)
        h = _poly_from_roots(N, roots)
        score = _phase_nonlinearity(h)
        # canonical orientation tie-break between a filter and its mirror: handled after selection
        if best is None or score < best[0] - 1e-12:
            best = (score, h)
    return best[1]


def coiflets():
    """coifN tables solved by tools/gen_coiflets.py (slow: run separately), dec_lo as mpf."""
    f = ROOT / "tools" / "coiflet_tables.json"
    if not f.exists():
        return {}
    return {k: [mp.mpf(x) for x in v] for k, v in json.loads(f.read_text()).items()}


def fmt(c):
    return mp.nstr(c, 25, strip_zeros=False)


def main():
    tables = {}
    for N in range(1, 39):                      # pywt: db1 .. db38
        tables[f"db{N}"] = daubechies(N)
    tables["haar"] = tables["db1"]
    for N in range(2, 21):                      # pywt: sym2 .. sym20
        tables[f"sym{N}"] = symlet(N)
    tables.update(coiflets())                   # pywt: coif1 .. coif17
    out_json = {k: [float(c) for c in v] for k, v in tables.items()}
    (ROOT / "oracle" / "wavelet_tables.json").write_text(json.dumps(out_json, indent=0))
    lines = ['"""Wavelet decomposition low-pass tables (== pywt.Wavelet(name).dec_lo), float64.',
             '',
             'GENERATED by tools/gen_wavelets.py (mpmath, 80 digits) — do not edit.  PyWavelets is not a dependency of',
             'this package: the four filters of an orthogonal wavelet all derive from this one list.',
             '"""',
             '',
             'DEC_LO = {']
    for k, v in out_json.items():
        lines.append(f'    "{k}": (')
        for c in v:
            lines.append(f'        {c!r},')
        lines.append('    ),')
    lines.append('}')
    (ROOT / "image-preprocessing-pipeline_b200" / "pystripe" / "_wavelet_tables.py").write_text("\n".join(lines) + "\n")
    print({k: len(v) for k, v in out_json.items()})


if __name__ == "__main__":
    main()
