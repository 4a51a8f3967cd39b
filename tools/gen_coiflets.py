#!/usr/bin/env python3
"""Coiflet low-pass filters coif1..coif17 (6N taps) from their defining equations, mpmath at 250 digits.

PyWavelets is not available offline, so the tables `pywt.Wavelet('coifN')` would return are
recomputed: unknown rec_lo h[0..6N-1] with origin tap 2N, consistent over-determined system
    sum_k h_k h_{k+2m} = delta_m                m = 0 .. 3N-1     (double-shift orthonormality)
    sum_k (-1)^k (k-2N)^p h_k = 0               p = 0 .. 2N-1     (2N vanishing wavelet moments)
    sum_k (k-2N)^p h_k = 0                      p = 1 .. 2N-1     (vanishing scaling moments; N-1 of them redundant)
solved in the least-squares sense (Gauss-Newton) by continuation in N (the order N-1 solution, re-centred, starts order N).  The branch this continuation follows
reproduces PyWavelets' published coif1, coif2, coif3 tables to the digits available (known-answer tests in
tests/test_oracle.py); for N >= 4 agreement with PyWavelets' tabulated branch is UNVERIFIED offline — what is verified
is that every table satisfies the defining equations to 1e-100.

Writes tools/coiflet_tables.json ({"coifN": dec_lo as decimal strings}); tools/gen_wavelets.py merges it.
"""
import json
import sys
import time
from pathlib import Path

import mpmath as mp

mp.mp.dps = 250
ROOT = Path(__file__).resolve().parent.parent


def solve(N, h0, max_iter=600):
    F = 6 * N
    c = 2 * N
    ks = [mp.mpf(k - c) for k in range(F)]
    sc = mp.mpf(2 * N)
    lin_rows = []
    for p in range(2 * N):
        lin_rows.append([((-1) ** k) * (ks[k] / sc) ** p for k in range(F)])
    for p in range(1, 2 * N):
        lin_rows.append([(ks[k] / sc) ** p for k in range(F)])
    h = [mp.mpf(x) for x in h0]

    def resid(hh):
        r = [mp.fsum(hh[k] * hh[k + 2 * m] for k in range(F - 2 * m)) - (1 if m == 0 else 0) for m in range(F // 2)]
        r += [mp.fsum(a * b for a, b in zip(row, hh)) for row in lin_rows]
        return r

    nrm = None
    for it in range(max_iter):
        r = resid(h)
        nrm = mp.sqrt(mp.fsum(x * x for x in r))
        if nrm < mp.mpf(10) ** (-100):
            break
        J = []
        for m in range(F // 2):
            row = [mp.mpf(0)] * F
            for i in range(F):
                if i + 2 * m < F:
                    row[i] += h[i + 2 * m]
                if i - 2 * m >= 0:
                    row[i] += h[i - 2 * m]
            J.append(row)
        J += lin_rows
        # least squares (the system is consistent but redundant): normal equations with a vanishing Levenberg term; the
        # working precision (250 digits) absorbs the squared condition number
        Jm = mp.matrix(J)
        A = Jm.T * Jm
        for i in range(F):
            A[i, i] += mp.mpf(10) ** (-150)
        dh = mp.lu_solve(A, -(Jm.T * mp.matrix(r)))
        step = mp.mpf(1)
        hn = h
        while step > mp.mpf(1) / 4096:
            hn = [h[i] + step * dh[i] for i in range(F)]
            rn = resid(hn)
            if mp.sqrt(mp.fsum(x * x for x in rn)) < nrm:
                break
            step /= 2
        h = hn
    return h, nrm, it


def main():
    n_max = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    # order 1 start: a windowed sinc centred on tap 2
    h = [mp.sinc(mp.pi * mp.mpf(k - 2) / 2) * mp.e ** (-(mp.mpf(k - 2) / mp.mpf("2.5")) ** 2) for k in range(6)]
    s = mp.fsum(h)
    h = [x * mp.sqrt(2) / s for x in h]
    out = {}
    f = ROOT / "tools" / "coiflet_tables.json"
    start = 1
    if f.exists():                              # resume after the highest order already solved
        out = json.loads(f.read_text())
        start = max(int(k[4:]) for k in out) + 1
        h = [mp.mpf(x) for x in out[f"coif{start - 1}"]][::-1]
    for N in range(start, n_max + 1):
        t = time.time()
        if N > 1:
            h = [mp.mpf(0), mp.mpf(0)] + list(h) + [mp.mpf(0)] * 4
        h, nrm, it = solve(N, h)
        if nrm > mp.mpf(10) ** (-90):
            print(f"coif{N}: Newton did not converge (residual {mp.nstr(nrm, 5)}); stopping", flush=True)
            break
        out[f"coif{N}"] = [mp.nstr(x, 40) for x in h[::-1]]      # dec_lo = reverse(rec_lo)
        print(f"coif{N}: residual {mp.nstr(nrm, 5)} after {it} iterations, peak tap {max(range(6 * N), key=lambda i: h[i])}, "
              f"sum {mp.nstr(mp.fsum(h), 20)}, {time.time() - t:.1f} s", flush=True)
        (ROOT / "tools" / "coiflet_tables.json").write_text(json.dumps(out, indent=0))


if __name__ == "__main__":
    main()
