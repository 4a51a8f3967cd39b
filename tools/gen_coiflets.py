#!/usr/bin/env python3
"""Coiflet low-pass filters coif1..coif17 (6N taps) from their defining equations, mpmath at 250 digits.

PyWavelets is not available offline, so the tables `pywt.Wavelet('coifN')` would return are
recomputed: unknown rec_lo h[0..6N-1] with origin tap 2N, consistent over-determined system
    sum_k h_k h_{k+2m} = delta_m                m = 0 .. 3N-1     (double-shift orthonormality)
    sum_k (-1)^k (k-2N)^p h_k = 0               p = 0 .. 2N-1     (2N vanishing wavelet moments)
    sum_k (k-2N)^p h_k = 0                      p = 1 .. 2N-1     (vanishing scaling moments; N-1 of them redundant)
solved in the least-squares sense (Gauss-Newton) by continuation in N (the order N-1 solution, re-centred, starts order N).  The branch this continuation follows
reproduces PyWavelets' published coif1, coif2, coif3 tables to the digits available (known-answer tests in
tests/test_oracle.py); for N >= 4 agreement with PyWavelets' tabulated branch is UNVERIFIED offline.  coif1..coif8 satisfy
the defining equations to 1e-100 (250-digit Newton); for N >= 7 the Jacobian is nearly singular at the solution (Newton
converges sub-linearly: 68 and 308 iterations for N = 7, 8), so coif9..coif17 are taken from a float64 Gauss-Newton in a
well-conditioned (Legendre) moment basis: orthonormality and moment residuals < 1e-13, taps determined to ~1e-5 or
worse along the flat direction.  The transforms use float32 taps, so perfect reconstruction is unaffected.

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


def float64_continuation(n_max):
    """Gauss-Newton in float64 with the moment conditions written in a Legendre basis (well conditioned):
    sum_k (-1)^k L_p(x_k) h_k = 0 and sum_k L_p(x_k) h_k = sqrt(2) L_p(x_c), p < 2N, x = taps mapped to [-1, 1].
    Converges for every N <= 17 to residuals ~1e-14 in a few seconds; used as the table itself where the 250-digit
    polish is impractical (N >= 9: the Jacobian is nearly singular at the solution, Newton converges sub-linearly)."""
    import numpy as np
    from numpy.polynomial import legendre as L

    def build(N):
        F, c = 6 * N, 2 * N
        k = np.arange(F)
        x = (k - (F - 1) / 2) / ((F - 1) / 2)
        xc = (c - (F - 1) / 2) / ((F - 1) / 2)
        rows, rhs = [], []
        for p in range(2 * N):
            co = np.zeros(p + 1); co[p] = 1
            rows.append(((-1.0) ** k) * L.legval(x, co)); rhs.append(0.0)
        for p in range(2 * N):
            co = np.zeros(p + 1); co[p] = 1
            rows.append(L.legval(x, co)); rhs.append(L.legval(xc, co) * np.sqrt(2))
        return np.array(rows), np.array(rhs)

    def resid(h, A, b):
        F = len(h)
        r = [np.dot(h[:F - 2 * m], h[2 * m:]) - (1.0 if m == 0 else 0.0) for m in range(F // 2)]
        return np.concatenate([r, A @ h - b])

    def jac(h, A):
        F = len(h)
        J = np.zeros((F // 2, F))
        for m in range(F // 2):
            J[m, :F - 2 * m] += h[2 * m:]
            J[m, 2 * m:] += h[:F - 2 * m]
        return np.vstack([J, A])

    k = np.arange(6) - 2
    h = np.sinc(k / 2.0) * np.exp(-(k / 2.5) ** 2)
    h *= np.sqrt(2) / h.sum()
    out = {}
    for N in range(1, n_max + 1):
        if N > 1:
            h = np.concatenate([[0, 0], h, [0, 0, 0, 0]])
        A, b = build(N)
        for _ in range(500):
            r = resid(h, A, b)
            nr = np.linalg.norm(r)
            if nr < 1e-14:
                break
            dh = np.linalg.lstsq(jac(h, A), -r, rcond=None)[0]
            st = 1.0
            while st > 1e-6 and np.linalg.norm(resid(h + st * dh, A, b)) >= nr:
                st /= 2
            h = h + st * dh
        out[N] = (h.copy(), float(nr))
        print(f"coif{N} (float64): residual {nr:.2e}, peak tap {int(np.argmax(h))}", flush=True)
    return out


def main():
    n_max = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    mp_max = int(sys.argv[2]) if len(sys.argv) > 2 else 6     # orders polished at 250 digits (7, 8 take minutes each)
    f = ROOT / "tools" / "coiflet_tables.json"
    out = json.loads(f.read_text()) if f.exists() else {}
    f64 = float64_continuation(n_max)
    for N in range(1, n_max + 1):
        name = f"coif{N}"
        if name in out and not out[name][0].endswith("f64"):
            continue                                          # already solved at high precision
        h64, _ = f64[N]
        if N <= mp_max:
            h, nrm, it = solve(N, [mp.mpf(float(x)) for x in h64], max_iter=80)
            if nrm < mp.mpf(10) ** (-90):
                out[name] = [mp.nstr(x, 40) for x in h[::-1]]
                print(f"{name}: polished, residual {mp.nstr(nrm, 5)} after {it} iterations", flush=True)
                continue
        out[name] = [repr(float(x)) for x in h64[::-1]]         # dec_lo = reverse(rec_lo), float64 accuracy
    f.write_text(json.dumps(out, indent=0))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
