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


# published PyWavelets dec_lo tables used to pin the root subset and the orientation of the low orders
SYM_KAT = {
    2: [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    3: [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388,
        0.3326705529509569],
    4: [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161, 0.29785779560527736,
        -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
    5: [0.027333068345077982, 0.029519490925774643, -0.039134249302383094, 0.1993975339773936, 0.7234076904024206,
        0.6339789634582119, 0.01660210576452232, -0.17532808990845047, -0.021101834024758855, 0.019538882735286728],
    6: [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148],
    7: [0.002681814568257878, -0.0010473848886829163, -0.01263630340325193, 0.03051551316596357, 0.0678926935013727,
        -0.049552834937127255, 0.017441255086855827, 0.5361019170917628, 0.767764317003164, 0.2886296317515146,
        -0.14004724044296152, -0.10780823770381774, 0.004010244871533663, 0.010268176708511255],
    8: [-0.0033824159510061256, -0.0005421323317911481, 0.03169508781149298, 0.007607487324917605, -0.1432942383508097,
        -0.061273359067658524, 0.4813596512583722, 0.7771857517005235, 0.3644418948353314, -0.05194583810770904,
        -0.027219029917056003, 0.049137179673607506, 0.003808752013890615, -0.01495225833704823,
        -0.0003029205147213668, 0.0018899503327594609],
}


def symlet(N):
    """Symlet: one inside/outside choice per conjugate root group of the Daubechies half-band polynomial.
    N <= 8: the subset and orientation that reproduce the published PyWavelets table (SYM_KAT; sym7 is NOT the subset a
    phase-linearity score picks).  N >= 9: least phase non-linearity, orientation with the centre of mass below the
    middle tap as in sym4/6/8 — agreement with PyWavelets UNVERIFIED offline."""
    pairs = _halfband_roots(N)
    groups = _group_conjugates(pairs)
    cands = []
    for mask in range(1 << len(groups)):
        roots = []
        for g, grp in enumerate(groups):
            pick = (mask >> g) & 1
            roots.extend(p[pick] for p in grp)
        cands.append(_poly_from_roots(N, roots))
    if N in SYM_KAT:
        ref = [mp.mpf(v) for v in SYM_KAT[N]]
        best = None
        for h in cands:
            for hh in (h, h[::-1]):
                d = max(abs(a - b) for a, b in zip(hh, ref))
                if best is None or d < best[0]:
                    best = (d, hh)
        assert best[0] < mp.mpf(10) ** (-9), (N, best[0])
        return best[1]
    h = min(cands, key=_phase_nonlinearity)
    F = len(h)
    com = mp.fsum(k * h[k] * h[k] for k in range(F)) / mp.fsum(x * x for x in h)
    return h if com < mp.mpf(F - 1) / 2 else h[::-1]


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
