/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under image-preprocessing-pipeline_b200/ may link or call this.
 *
 * CPU restatement (float32) of the third-party FFT the reference's hot path executes:
 *   scipy.fftpack.rfft / irfft  (call sites /root/reference/pystripe/core.py:751,753)
 *   -> scipy.fft._pocketfft.pypocketfft.r2r_fftpack -> pocketfft_hdronly.hpp (scipy 1.18.1 vendored copy; the
 *      reference does not pin scipy): rfftp<float> (radf2/3/4/5/g, radb2/3/4/5/g), and for lengths where
 *      pocketfft's cost model prefers it, fftblue<float> over cfftp<float> (pass2/3/4/5/7/8/11).
 * The algorithm is restated from the published source; operation ORDER matters here because the GPU kernels mirror it
 * rounding for rounding.  PINNED: tests/test_oracle.py compares this file BIT FOR BIT with scipy.fftpack.rfft / irfft
 * and scipy.fft.fft (complex64) of the scipy installed in the image, for every length 1..N and random data.
 *
 * Build: oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, like the x86-64 baseline scipy wheel).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float r, i; } cf;

/* ---------------------------------------------------------------- sincos_2pibyn<float> (Thigh = double) */
typedef struct { size_t N, mask, shift; double *v1, *v2; } sincos_t;

static void sc_calc(size_t x, size_t n, double ang, double *res)
{
    x <<= 3;
    if (x < 4 * n) {
        if (x < 2 * n) {
            if (x < n) { res[0] = cos((double)x * ang); res[1] = sin((double)x * ang); return; }
            res[0] = sin((double)(2 * n - x) * ang); res[1] = cos((double)(2 * n - x) * ang); return;
        } else {
            x -= 2 * n;
            if (x < n) { res[0] = -sin((double)x * ang); res[1] = cos((double)x * ang); return; }
            res[0] = -cos((double)(2 * n - x) * ang); res[1] = sin((double)(2 * n - x) * ang); return;
        }
    } else {
        x = 8 * n - x;
        if (x < 2 * n) {
            if (x < n) { res[0] = cos((double)x * ang); res[1] = -sin((double)x * ang); return; }
            res[0] = sin((double)(2 * n - x) * ang); res[1] = -cos((double)(2 * n - x) * ang); return;
        } else {
            x -= 2 * n;
            if (x < n) { res[0] = -sin((double)x * ang); res[1] = -cos((double)x * ang); return; }
            res[0] = -cos((double)(2 * n - x) * ang); res[1] = -sin((double)(2 * n - x) * ang); return;
        }
    }
}

static void sc_init(sincos_t *s, size_t n)
{
    const long double pi = 3.141592653589793238462643383279502884197L;
    double ang = (double)(0.25L * pi / (long double)n);
    size_t nval = (n + 2) / 2;
    s->N = n;
    s->shift = 1;
    while (((size_t)1 << s->shift) * ((size_t)1 << s->shift) < nval) ++s->shift;
    s->mask = ((size_t)1 << s->shift) - 1;
    size_t n1 = s->mask + 1, n2 = (nval + s->mask) / (s->mask + 1);
    s->v1 = (double *)malloc(sizeof(double) * 2 * n1);
    s->v2 = (double *)malloc(sizeof(double) * 2 * n2);
    s->v1[0] = 1.0; s->v1[1] = 0.0;
    for (size_t i = 1; i < n1; ++i) sc_calc(i, n, ang, s->v1 + 2 * i);
    s->v2[0] = 1.0; s->v2[1] = 0.0;
    for (size_t i = 1; i < n2; ++i) sc_calc(i * (s->mask + 1), n, ang, s->v2 + 2 * i);
}
static void sc_free(sincos_t *s) { free(s->v1); free(s->v2); }
static cf sc_get(const sincos_t *s, size_t idx)
{
    cf r;
    if (2 * idx <= s->N) {
        const double *x1 = s->v1 + 2 * (idx & s->mask), *x2 = s->v2 + 2 * (idx >> s->shift);
        r.r = (float)(x1[0] * x2[0] - x1[1] * x2[1]);
        r.i = (float)(x1[0] * x2[1] + x1[1] * x2[0]);
        return r;
    }
    idx = s->N - idx;
    const double *x1 = s->v1 + 2 * (idx & s->mask), *x2 = s->v2 + 2 * (idx >> s->shift);
    r.r = (float)(x1[0] * x2[0] - x1[1] * x2[1]);
    r.i = -(float)(x1[0] * x2[1] + x1[1] * x2[0]);
    return r;
}

/* exported so the GPU host code's tables can be compared with these in tests */
void orc_pf_twiddle(size_t n, size_t idx, float *re, float *im)
{
    sincos_t s; sc_init(&s, n);
    cf v = sc_get(&s, idx);
    *re = v.r; *im = v.i;
    sc_free(&s);
}

/* ---------------------------------------------------------------- util */
__attribute__((unused)) static size_t largest_prime_factor(size_t n)
{
    size_t res = 1;
    while ((n & 1) == 0) { res = 2; n >>= 1; }
    for (size_t x = 3; x * x <= n; x += 2)
        while ((n % x) == 0) { res = x; n /= x; }
    if (n > 1) res = n;
    return res;
}
__attribute__((unused)) static double cost_guess(size_t n)
{
    const double lfp = 1.1;
    size_t ni = n;
    double result = 0.;
    while ((n & 3) == 0) { result += 2; n >>= 2; }
    while ((n & 1) == 0) { result += 1; n >>= 1; }
    for (size_t x = 3; x * x <= n; x += 2)
        while ((n % x) == 0) { result += (x <= 5) ? (double)x : lfp * (double)x; n /= x; }
    if (n > 1) result += (n <= 5) ? (double)n : lfp * (double)n;
    return result * (double)ni;
}
/* smallest 11-smooth number >= n */
static size_t good_size_cmplx(size_t n)
{
    if (n <= 12) return n;
    size_t bestfac = 2 * n;
    for (size_t f11 = 1; f11 < bestfac; f11 *= 11)
        for (size_t f117 = f11; f117 < bestfac; f117 *= 7)
            for (size_t f1175 = f117; f1175 < bestfac; f1175 *= 5) {
                size_t x = f1175;
                while (x < n) x *= 2;
                for (;;) {
                    if (x < n) x *= 3;
                    else if (x > n) {
                        if (x < bestfac) bestfac = x;
                        if (x & 1) break;
                        x >>= 1;
                    } else return n;
                }
            }
    return bestfac;
}
size_t orc_pf_good_size(size_t n) { return good_size_cmplx(n); }

/* ================================================================ rfftp<float> */
#define MAXFACT 32
typedef struct { size_t fct; float *tw, *tws; } rfct;
typedef struct { size_t length, nfct; rfct fct[MAXFACT]; float *mem; } rfftp_t;

static size_t orc_ovr_r[MAXFACT], orc_ovr_rn = 0, orc_ovr_c[MAXFACT], orc_ovr_cn = 0;
void orc_override_factors(int real, const size_t *f, size_t nf) { if (real) { orc_ovr_rn = nf; for (size_t i = 0; i < nf; ++i) orc_ovr_r[i] = f[i]; } else { orc_ovr_cn = nf; for (size_t i = 0; i < nf; ++i) orc_ovr_c[i] = f[i]; } }
static void rfftp_init(rfftp_t *p, size_t length)
{
    p->length = length;
    p->nfct = 0;
    p->mem = NULL;
    if (length == 1) return;
    size_t len = length;
    while ((len % 4) == 0) { p->fct[p->nfct++].fct = 4; len >>= 2; }
    if ((len % 2) == 0) {
        len >>= 1;
        p->fct[p->nfct++].fct = 2;
        size_t t = p->fct[0].fct; p->fct[0].fct = p->fct[p->nfct - 1].fct; p->fct[p->nfct - 1].fct = t;
    }
    for (size_t divisor = 3; divisor * divisor <= len; divisor += 2)
        while ((len % divisor) == 0) { p->fct[p->nfct++].fct = divisor; len /= divisor; }
    if (len > 1) p->fct[p->nfct++].fct = len;
    if (orc_ovr_rn) { p->nfct = orc_ovr_rn; for (size_t i = 0; i < orc_ovr_rn; ++i) p->fct[i].fct = orc_ovr_r[i]; }
    /* twiddles */
    size_t twsz = 0, l1 = 1;
    for (size_t k = 0; k < p->nfct; ++k) {
        size_t ip = p->fct[k].fct, ido = length / (l1 * ip);
        twsz += (ip - 1) * (ido - 1);
        if (ip > 5) twsz += 2 * ip;
        l1 *= ip;
    }
    p->mem = (float *)calloc(twsz + 1, sizeof(float));
    sincos_t twid; sc_init(&twid, length);
    l1 = 1;
    float *ptr = p->mem;
    for (size_t k = 0; k < p->nfct; ++k) {
        size_t ip = p->fct[k].fct, ido = length / (l1 * ip);
        p->fct[k].tw = p->fct[k].tws = NULL;
        if (k < p->nfct - 1) {
            p->fct[k].tw = ptr; ptr += (ip - 1) * (ido - 1);
            for (size_t j = 1; j < ip; ++j)
                for (size_t i = 1; i <= (ido - 1) / 2; ++i) {
                    cf t = sc_get(&twid, j * l1 * i);
                    p->fct[k].tw[(j - 1) * (ido - 1) + 2 * i - 2] = t.r;
                    p->fct[k].tw[(j - 1) * (ido - 1) + 2 * i - 1] = t.i;
                }
        }
        if (ip > 5) {
            p->fct[k].tws = ptr; ptr += 2 * ip;
            p->fct[k].tws[0] = 1.f;
            p->fct[k].tws[1] = 0.f;
            for (size_t i = 2, ic = 2 * ip - 2; i <= ic; i += 2, ic -= 2) {
                cf t = sc_get(&twid, i / 2 * (length / ip));
                p->fct[k].tws[i] = t.r;
                p->fct[k].tws[i + 1] = t.i;
                p->fct[k].tws[ic] = t.r;
                p->fct[k].tws[ic + 1] = -t.i;
            }
        }
        l1 *= ip;
    }
    sc_free(&twid);
}
static void rfftp_free(rfftp_t *p) { free(p->mem); }

#define PM(a, b, c, d) { a = (c) + (d); b = (c) - (d); }
/* (a+ib) = conj(c+id) * (e+if) */
#define MULPM(a, b, c, d, e, f) { a = (c) * (e) + (d) * (f); b = (c) * (f) - (d) * (e); }
#define WA(x, i) wa[(i) + (x) * (ido - 1)]

static void radf2(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
#define CC(a, b, c) cc[(a) + ido * ((b) + l1 * (c))]
#define CH(a, b, c) ch[(a) + ido * ((b) + 2 * (c))]
    for (size_t k = 0; k < l1; k++) PM(CH(0, 0, k), CH(ido - 1, 1, k), CC(0, k, 0), CC(0, k, 1))
    if ((ido & 1) == 0)
        for (size_t k = 0; k < l1; k++) {
            CH(0, 1, k) = -CC(ido - 1, k, 1);
            CH(ido - 1, 0, k) = CC(ido - 1, k, 0);
        }
    if (ido <= 2) return;
    for (size_t k = 0; k < l1; k++)
        for (size_t i = 2; i < ido; i += 2) {
            size_t ic = ido - i;
            float tr2, ti2;
            MULPM(tr2, ti2, WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            PM(CH(i - 1, 0, k), CH(ic - 1, 1, k), CC(i - 1, k, 0), tr2)
            PM(CH(i, 0, k), CH(ic, 1, k), ti2, CC(i, k, 0))
        }
#undef CH
}

static void radf3(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float taur = -0.5f, taui = (float)0.8660254037844386467637231707529362L;
#define CH(a, b, c) ch[(a) + ido * ((b) + 3 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float cr2 = CC(0, k, 1) + CC(0, k, 2);
        CH(0, 0, k) = CC(0, k, 0) + cr2;
        CH(0, 2, k) = taui * (CC(0, k, 2) - CC(0, k, 1));
        CH(ido - 1, 1, k) = CC(0, k, 0) + taur * cr2;
    }
    if (ido == 1) return;
    for (size_t k = 0; k < l1; k++)
        for (size_t i = 2; i < ido; i += 2) {
            size_t ic = ido - i;
            float ci2, di2, di3, cr2, dr2, dr3, ti2, ti3, tr2, tr3;
            MULPM(dr2, di2, WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            MULPM(dr3, di3, WA(1, i - 2), WA(1, i - 1), CC(i - 1, k, 2), CC(i, k, 2))
            cr2 = dr2 + dr3;
            ci2 = di2 + di3;
            CH(i - 1, 0, k) = CC(i - 1, k, 0) + cr2;
            CH(i, 0, k) = CC(i, k, 0) + ci2;
            tr2 = CC(i - 1, k, 0) + taur * cr2;
            ti2 = CC(i, k, 0) + taur * ci2;
            tr3 = taui * (di2 - di3);
            ti3 = taui * (dr3 - dr2);
            PM(CH(i - 1, 2, k), CH(ic - 1, 1, k), tr2, tr3)
            PM(CH(i, 2, k), CH(ic, 1, k), ti3, ti2)
        }
#undef CH
}

static void radf4(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float hsqt2 = (float)0.707106781186547524400844362104849L;
#define CH(a, b, c) ch[(a) + ido * ((b) + 4 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float tr1, tr2;
        PM(tr1, CH(0, 2, k), CC(0, k, 3), CC(0, k, 1))
        PM(tr2, CH(ido - 1, 1, k), CC(0, k, 0), CC(0, k, 2))
        PM(CH(0, 0, k), CH(ido - 1, 3, k), tr2, tr1)
    }
    if ((ido & 1) == 0)
        for (size_t k = 0; k < l1; k++) {
            float ti1 = -hsqt2 * (CC(ido - 1, k, 1) + CC(ido - 1, k, 3));
            float tr1 = hsqt2 * (CC(ido - 1, k, 1) - CC(ido - 1, k, 3));
            PM(CH(ido - 1, 0, k), CH(ido - 1, 2, k), CC(ido - 1, k, 0), tr1)
            PM(CH(0, 3, k), CH(0, 1, k), ti1, CC(ido - 1, k, 2))
        }
    if (ido <= 2) return;
    for (size_t k = 0; k < l1; k++)
        for (size_t i = 2; i < ido; i += 2) {
            size_t ic = ido - i;
            float ci2, ci3, ci4, cr2, cr3, cr4, ti1, ti2, ti3, ti4, tr1, tr2, tr3, tr4;
            MULPM(cr2, ci2, WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            MULPM(cr3, ci3, WA(1, i - 2), WA(1, i - 1), CC(i - 1, k, 2), CC(i, k, 2))
            MULPM(cr4, ci4, WA(2, i - 2), WA(2, i - 1), CC(i - 1, k, 3), CC(i, k, 3))
            PM(tr1, tr4, cr4, cr2)
            PM(ti1, ti4, ci2, ci4)
            PM(tr2, tr3, CC(i - 1, k, 0), cr3)
            PM(ti2, ti3, CC(i, k, 0), ci3)
            PM(CH(i - 1, 0, k), CH(ic - 1, 3, k), tr2, tr1)
            PM(CH(i, 0, k), CH(ic, 3, k), ti1, ti2)
            PM(CH(i - 1, 2, k), CH(ic - 1, 1, k), tr3, ti4)
            PM(CH(i, 2, k), CH(ic, 1, k), tr4, ti3)
        }
#undef CH
}

static void radf5(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float tr11 = (float)0.3090169943749474241022934171828191L, ti11 = (float)0.9510565162951535721164393333793821L,
                tr12 = (float)-0.8090169943749474241022934171828191L, ti12 = (float)0.5877852522924731291687059546390728L;
#define CH(a, b, c) ch[(a) + ido * ((b) + 5 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float cr2, cr3, ci4, ci5;
        PM(cr2, ci5, CC(0, k, 4), CC(0, k, 1))
        PM(cr3, ci4, CC(0, k, 3), CC(0, k, 2))
        CH(0, 0, k) = CC(0, k, 0) + cr2 + cr3;
        CH(ido - 1, 1, k) = CC(0, k, 0) + tr11 * cr2 + tr12 * cr3;
        CH(0, 2, k) = ti11 * ci5 + ti12 * ci4;
        CH(ido - 1, 3, k) = CC(0, k, 0) + tr12 * cr2 + tr11 * cr3;
        CH(0, 4, k) = ti12 * ci5 - ti11 * ci4;
    }
    if (ido == 1) return;
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 2, ic = ido - 2; i < ido; i += 2, ic -= 2) {
            float di2, di3, di4, di5, dr2, dr3, dr4, dr5;
            MULPM(dr2, di2, WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            MULPM(dr3, di3, WA(1, i - 2), WA(1, i - 1), CC(i - 1, k, 2), CC(i, k, 2))
            MULPM(dr4, di4, WA(2, i - 2), WA(2, i - 1), CC(i - 1, k, 3), CC(i, k, 3))
            MULPM(dr5, di5, WA(3, i - 2), WA(3, i - 1), CC(i - 1, k, 4), CC(i, k, 4))
            float cr2, cr3, cr4, cr5, ci2, ci3, ci4, ci5;
            PM(cr2, ci5, dr5, dr2)
            PM(ci2, cr5, di2, di5)
            PM(cr3, ci4, dr4, dr3)
            PM(ci3, cr4, di3, di4)
            CH(i - 1, 0, k) = CC(i - 1, k, 0) + cr2 + cr3;
            CH(i, 0, k) = CC(i, k, 0) + ci2 + ci3;
            float tr2 = CC(i - 1, k, 0) + tr11 * cr2 + tr12 * cr3;
            float ti2 = CC(i, k, 0) + tr11 * ci2 + tr12 * ci3;
            float tr3 = CC(i - 1, k, 0) + tr12 * cr2 + tr11 * cr3;
            float ti3 = CC(i, k, 0) + tr12 * ci2 + tr11 * ci3;
            float tr5 = cr5 * ti11 + cr4 * ti12;
            float ti5 = ci5 * ti11 + ci4 * ti12;
            float tr4 = cr5 * ti12 - cr4 * ti11;
            float ti4 = ci5 * ti12 - ci4 * ti11;
            PM(CH(i - 1, 2, k), CH(ic - 1, 1, k), tr2, tr5)
            PM(CH(i, 2, k), CH(ic, 1, k), ti5, ti2)
            PM(CH(i - 1, 4, k), CH(ic - 1, 3, k), tr3, tr4)
            PM(CH(i, 4, k), CH(ic, 3, k), ti4, ti3)
        }
#undef CH
#undef CC
}

/* generic odd radix; result ends in cc (the caller swaps twice) */
static void radfg(size_t ido, size_t ip, size_t l1, float *cc, float *ch, const float *wa, const float *csarr)
{
    const size_t cdim = ip;
    size_t ipph = (ip + 1) / 2;
    size_t idl1 = ido * l1;
#define CC(a, b, c) cc[(a) + ido * ((b) + cdim * (c))]
#define CH(a, b, c) ch[(a) + ido * ((b) + l1 * (c))]
#define C1(a, b, c) cc[(a) + ido * ((b) + l1 * (c))]
#define C2(a, b) cc[(a) + idl1 * (b)]
#define CH2(a, b) ch[(a) + idl1 * (b)]
    if (ido > 1) {
        for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc) {
            size_t is = (j - 1) * (ido - 1), is2 = (jc - 1) * (ido - 1);
            for (size_t k = 0; k < l1; ++k) {
                size_t idij = is, idij2 = is2;
                for (size_t i = 1; i <= ido - 2; i += 2) {
                    float t1 = C1(i, k, j), t2 = C1(i + 1, k, j), t3 = C1(i, k, jc), t4 = C1(i + 1, k, jc);
                    float x1 = wa[idij] * t1 + wa[idij + 1] * t2, x2 = wa[idij] * t2 - wa[idij + 1] * t1,
                          x3 = wa[idij2] * t3 + wa[idij2 + 1] * t4, x4 = wa[idij2] * t4 - wa[idij2 + 1] * t3;
                    PM(C1(i, k, j), C1(i + 1, k, jc), x3, x1)
                    PM(C1(i + 1, k, j), C1(i, k, jc), x2, x4)
                    idij += 2;
                    idij2 += 2;
                }
            }
        }
    }
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
        for (size_t k = 0; k < l1; ++k) {
            float t1 = C1(0, k, j), t2 = C1(0, k, jc);
            PM(C1(0, k, j), C1(0, k, jc), t2, t1)
        }
    for (size_t l = 1, lc = ip - 1; l < ipph; ++l, --lc) {
        for (size_t ik = 0; ik < idl1; ++ik) {
            CH2(ik, l) = C2(ik, 0) + csarr[2 * l] * C2(ik, 1) + csarr[4 * l] * C2(ik, 2);
            CH2(ik, lc) = csarr[2 * l + 1] * C2(ik, ip - 1) + csarr[4 * l + 1] * C2(ik, ip - 2);
        }
        size_t iang = 2 * l;
        size_t j = 3, jc = ip - 3;
        for (; j + 3 < ipph; j += 4, jc -= 4) {
            iang += l; if (iang >= ip) iang -= ip;
            float ar1 = csarr[2 * iang], ai1 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar2 = csarr[2 * iang], ai2 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar3 = csarr[2 * iang], ai3 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar4 = csarr[2 * iang], ai4 = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                CH2(ik, l) += ar1 * C2(ik, j) + ar2 * C2(ik, j + 1) + ar3 * C2(ik, j + 2) + ar4 * C2(ik, j + 3);
                CH2(ik, lc) += ai1 * C2(ik, jc) + ai2 * C2(ik, jc - 1) + ai3 * C2(ik, jc - 2) + ai4 * C2(ik, jc - 3);
            }
        }
        for (; j + 1 < ipph; j += 2, jc -= 2) {
            iang += l; if (iang >= ip) iang -= ip;
            float ar1 = csarr[2 * iang], ai1 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar2 = csarr[2 * iang], ai2 = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                CH2(ik, l) += ar1 * C2(ik, j) + ar2 * C2(ik, j + 1);
                CH2(ik, lc) += ai1 * C2(ik, jc) + ai2 * C2(ik, jc - 1);
            }
        }
        for (; j < ipph; ++j, --jc) {
            iang += l; if (iang >= ip) iang -= ip;
            float ar = csarr[2 * iang], ai = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                CH2(ik, l) += ar * C2(ik, j);
                CH2(ik, lc) += ai * C2(ik, jc);
            }
        }
    }
    for (size_t ik = 0; ik < idl1; ++ik) CH2(ik, 0) = C2(ik, 0);
    for (size_t j = 1; j < ipph; ++j)
        for (size_t ik = 0; ik < idl1; ++ik) CH2(ik, 0) += C2(ik, j);
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) CC(i, 0, k) = CH(i, k, 0);
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc) {
        size_t j2 = 2 * j - 1;
        for (size_t k = 0; k < l1; ++k) {
            CC(ido - 1, j2, k) = CH(0, k, j);
            CC(0, j2 + 1, k) = CH(0, k, jc);
        }
    }
    if (ido == 1) return;
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc) {
        size_t j2 = 2 * j - 1;
        for (size_t k = 0; k < l1; ++k)
            for (size_t i = 1, ic = ido - i - 2; i <= ido - 2; i += 2, ic -= 2) {
                CC(i, j2 + 1, k) = CH(i, k, j) + CH(i, k, jc);
                CC(ic, j2, k) = CH(i, k, j) - CH(i, k, jc);
                CC(i + 1, j2 + 1, k) = CH(i + 1, k, j) + CH(i + 1, k, jc);
                CC(ic + 1, j2, k) = CH(i + 1, k, jc) - CH(i + 1, k, j);
            }
    }
#undef CC
#undef CH
}

/* ---- backward */
static void radb2(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
#define CC(a, b, c) cc[(a) + ido * ((b) + 2 * (c))]
#define CH(a, b, c) ch[(a) + ido * ((b) + l1 * (c))]
    for (size_t k = 0; k < l1; k++) PM(CH(0, k, 0), CH(0, k, 1), CC(0, 0, k), CC(ido - 1, 1, k))
    if ((ido & 1) == 0)
        for (size_t k = 0; k < l1; k++) {
            CH(ido - 1, k, 0) = 2.f * CC(ido - 1, 0, k);
            CH(ido - 1, k, 1) = -2.f * CC(0, 1, k);
        }
    if (ido <= 2) return;
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 2; i < ido; i += 2) {
            size_t ic = ido - i;
            float ti2, tr2;
            PM(CH(i - 1, k, 0), tr2, CC(i - 1, 0, k), CC(ic - 1, 1, k))
            PM(ti2, CH(i, k, 0), CC(i, 0, k), CC(ic, 1, k))
            MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), ti2, tr2)
        }
#undef CC
}

static void radb3(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float taur = -0.5f, taui = (float)0.8660254037844386467637231707529362L;
#define CC(a, b, c) cc[(a) + ido * ((b) + 3 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float tr2 = 2.f * CC(ido - 1, 1, k);
        float cr2 = CC(0, 0, k) + taur * tr2;
        CH(0, k, 0) = CC(0, 0, k) + tr2;
        float ci3 = 2.f * taui * CC(0, 2, k);
        PM(CH(0, k, 2), CH(0, k, 1), cr2, ci3)
    }
    if (ido == 1) return;
    for (size_t k = 0; k < l1; k++)
        for (size_t i = 2, ic = ido - 2; i < ido; i += 2, ic -= 2) {
            float tr2 = CC(i - 1, 2, k) + CC(ic - 1, 1, k);
            float ti2 = CC(i, 2, k) - CC(ic, 1, k);
            float cr2 = CC(i - 1, 0, k) + taur * tr2;
            float ci2 = CC(i, 0, k) + taur * ti2;
            CH(i - 1, k, 0) = CC(i - 1, 0, k) + tr2;
            CH(i, k, 0) = CC(i, 0, k) + ti2;
            float cr3 = taui * (CC(i - 1, 2, k) - CC(ic - 1, 1, k));
            float ci3 = taui * (CC(i, 2, k) + CC(ic, 1, k));
            float di2, di3, dr2, dr3;
            PM(dr3, dr2, cr2, ci3)
            PM(di2, di3, ci2, cr3)
            MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), di2, dr2)
            MULPM(CH(i, k, 2), CH(i - 1, k, 2), WA(1, i - 2), WA(1, i - 1), di3, dr3)
        }
#undef CC
}

static void radb4(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float sqrt2 = (float)1.414213562373095048801688724209698L;
#define CC(a, b, c) cc[(a) + ido * ((b) + 4 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float tr1, tr2;
        PM(tr2, tr1, CC(0, 0, k), CC(ido - 1, 3, k))
        float tr3 = 2.f * CC(ido - 1, 1, k);
        float tr4 = 2.f * CC(0, 2, k);
        PM(CH(0, k, 0), CH(0, k, 2), tr2, tr3)
        PM(CH(0, k, 3), CH(0, k, 1), tr1, tr4)
    }
    if ((ido & 1) == 0)
        for (size_t k = 0; k < l1; k++) {
            float tr1, tr2, ti1, ti2;
            PM(ti1, ti2, CC(0, 3, k), CC(0, 1, k))
            PM(tr2, tr1, CC(ido - 1, 0, k), CC(ido - 1, 2, k))
            CH(ido - 1, k, 0) = tr2 + tr2;
            CH(ido - 1, k, 1) = sqrt2 * (tr1 - ti1);
            CH(ido - 1, k, 2) = ti2 + ti2;
            CH(ido - 1, k, 3) = -sqrt2 * (tr1 + ti1);
        }
    if (ido <= 2) return;
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 2; i < ido; i += 2) {
            float ci2, ci3, ci4, cr2, cr3, cr4, ti1, ti2, ti3, ti4, tr1, tr2, tr3, tr4;
            size_t ic = ido - i;
            PM(tr2, tr1, CC(i - 1, 0, k), CC(ic - 1, 3, k))
            PM(ti1, ti2, CC(i, 0, k), CC(ic, 3, k))
            PM(tr4, ti3, CC(i, 2, k), CC(ic, 1, k))
            PM(tr3, ti4, CC(i - 1, 2, k), CC(ic - 1, 1, k))
            PM(CH(i - 1, k, 0), cr3, tr2, tr3)
            PM(CH(i, k, 0), ci3, ti2, ti3)
            PM(cr4, cr2, tr1, tr4)
            PM(ci2, ci4, ti1, ti4)
            MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), ci2, cr2)
            MULPM(CH(i, k, 2), CH(i - 1, k, 2), WA(1, i - 2), WA(1, i - 1), ci3, cr3)
            MULPM(CH(i, k, 3), CH(i - 1, k, 3), WA(2, i - 2), WA(2, i - 1), ci4, cr4)
        }
#undef CC
}

static void radb5(size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    const float tr11 = (float)0.3090169943749474241022934171828191L, ti11 = (float)0.9510565162951535721164393333793821L,
                tr12 = (float)-0.8090169943749474241022934171828191L, ti12 = (float)0.5877852522924731291687059546390728L;
#define CC(a, b, c) cc[(a) + ido * ((b) + 5 * (c))]
    for (size_t k = 0; k < l1; k++) {
        float ti5 = CC(0, 2, k) + CC(0, 2, k);
        float ti4 = CC(0, 4, k) + CC(0, 4, k);
        float tr2 = CC(ido - 1, 1, k) + CC(ido - 1, 1, k);
        float tr3 = CC(ido - 1, 3, k) + CC(ido - 1, 3, k);
        CH(0, k, 0) = CC(0, 0, k) + tr2 + tr3;
        float cr2 = CC(0, 0, k) + tr11 * tr2 + tr12 * tr3;
        float cr3 = CC(0, 0, k) + tr12 * tr2 + tr11 * tr3;
        float ci4, ci5;
        MULPM(ci5, ci4, ti5, ti4, ti11, ti12)
        PM(CH(0, k, 4), CH(0, k, 1), cr2, ci5)
        PM(CH(0, k, 3), CH(0, k, 2), cr3, ci4)
    }
    if (ido == 1) return;
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 2, ic = ido - 2; i < ido; i += 2, ic -= 2) {
            float tr2, tr3, tr4, tr5, ti2, ti3, ti4, ti5;
            PM(tr2, tr5, CC(i - 1, 2, k), CC(ic - 1, 1, k))
            PM(ti5, ti2, CC(i, 2, k), CC(ic, 1, k))
            PM(tr3, tr4, CC(i - 1, 4, k), CC(ic - 1, 3, k))
            PM(ti4, ti3, CC(i, 4, k), CC(ic, 3, k))
            CH(i - 1, k, 0) = CC(i - 1, 0, k) + tr2 + tr3;
            CH(i, k, 0) = CC(i, 0, k) + ti2 + ti3;
            float cr2 = CC(i - 1, 0, k) + tr11 * tr2 + tr12 * tr3;
            float ci2 = CC(i, 0, k) + tr11 * ti2 + tr12 * ti3;
            float cr3 = CC(i - 1, 0, k) + tr12 * tr2 + tr11 * tr3;
            float ci3 = CC(i, 0, k) + tr12 * ti2 + tr11 * ti3;
            float ci4, ci5, cr5, cr4;
            MULPM(cr5, cr4, tr5, tr4, ti11, ti12)
            MULPM(ci5, ci4, ti5, ti4, ti11, ti12)
            float dr2, dr3, dr4, dr5, di2, di3, di4, di5;
            PM(dr4, dr3, cr3, ci4)
            PM(di3, di4, ci3, cr4)
            PM(dr5, dr2, cr2, ci5)
            PM(di2, di5, ci2, cr5)
            MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), di2, dr2)
            MULPM(CH(i, k, 2), CH(i - 1, k, 2), WA(1, i - 2), WA(1, i - 1), di3, dr3)
            MULPM(CH(i, k, 3), CH(i - 1, k, 3), WA(2, i - 2), WA(2, i - 1), di4, dr4)
            MULPM(CH(i, k, 4), CH(i - 1, k, 4), WA(3, i - 2), WA(3, i - 1), di5, dr5)
        }
#undef CC
#undef CH
}

static void radbg(size_t ido, size_t ip, size_t l1, float *cc, float *ch, const float *wa, const float *csarr)
{
    const size_t cdim = ip;
    size_t ipph = (ip + 1) / 2;
    size_t idl1 = ido * l1;
#define CC(a, b, c) cc[(a) + ido * ((b) + cdim * (c))]
#define CH(a, b, c) ch[(a) + ido * ((b) + l1 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) CH(i, k, 0) = CC(i, 0, k);
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc) {
        size_t j2 = 2 * j - 1;
        for (size_t k = 0; k < l1; ++k) {
            CH(0, k, j) = 2 * CC(ido - 1, j2, k);
            CH(0, k, jc) = 2 * CC(0, j2 + 1, k);
        }
    }
    if (ido != 1) {
        for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc) {
            size_t j2 = 2 * j - 1;
            for (size_t k = 0; k < l1; ++k)
                for (size_t i = 1, ic = ido - i - 2; i <= ido - 2; i += 2, ic -= 2) {
                    CH(i, k, j) = CC(i, j2 + 1, k) + CC(ic, j2, k);
                    CH(i, k, jc) = CC(i, j2 + 1, k) - CC(ic, j2, k);
                    CH(i + 1, k, j) = CC(i + 1, j2 + 1, k) - CC(ic + 1, j2, k);
                    CH(i + 1, k, jc) = CC(i + 1, j2 + 1, k) + CC(ic + 1, j2, k);
                }
        }
    }
    for (size_t l = 1, lc = ip - 1; l < ipph; ++l, --lc) {
        for (size_t ik = 0; ik < idl1; ++ik) {
            C2(ik, l) = CH2(ik, 0) + csarr[2 * l] * CH2(ik, 1) + csarr[4 * l] * CH2(ik, 2);
            C2(ik, lc) = csarr[2 * l + 1] * CH2(ik, ip - 1) + csarr[4 * l + 1] * CH2(ik, ip - 2);
        }
        size_t iang = 2 * l;
        size_t j = 3, jc = ip - 3;
        for (; j + 3 < ipph; j += 4, jc -= 4) {
            iang += l; if (iang >= ip) iang -= ip;
            float ar1 = csarr[2 * iang], ai1 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar2 = csarr[2 * iang], ai2 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar3 = csarr[2 * iang], ai3 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar4 = csarr[2 * iang], ai4 = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                C2(ik, l) += ar1 * CH2(ik, j) + ar2 * CH2(ik, j + 1) + ar3 * CH2(ik, j + 2) + ar4 * CH2(ik, j + 3);
                C2(ik, lc) += ai1 * CH2(ik, jc) + ai2 * CH2(ik, jc - 1) + ai3 * CH2(ik, jc - 2) + ai4 * CH2(ik, jc - 3);
            }
        }
        for (; j + 1 < ipph; j += 2, jc -= 2) {
            iang += l; if (iang >= ip) iang -= ip;
            float ar1 = csarr[2 * iang], ai1 = csarr[2 * iang + 1];
            iang += l; if (iang >= ip) iang -= ip;
            float ar2 = csarr[2 * iang], ai2 = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                C2(ik, l) += ar1 * CH2(ik, j) + ar2 * CH2(ik, j + 1);
                C2(ik, lc) += ai1 * CH2(ik, jc) + ai2 * CH2(ik, jc - 1);
            }
        }
        for (; j < ipph; ++j, --jc) {
            iang += l; if (iang >= ip) iang -= ip;
            float war = csarr[2 * iang], wai = csarr[2 * iang + 1];
            for (size_t ik = 0; ik < idl1; ++ik) {
                C2(ik, l) += war * CH2(ik, j);
                C2(ik, lc) += wai * CH2(ik, jc);
            }
        }
    }
    for (size_t j = 1; j < ipph; ++j)
        for (size_t ik = 0; ik < idl1; ++ik) CH2(ik, 0) += CH2(ik, j);
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
        for (size_t k = 0; k < l1; ++k) PM(CH(0, k, jc), CH(0, k, j), C1(0, k, j), C1(0, k, jc))
    if (ido == 1) return;
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
        for (size_t k = 0; k < l1; ++k)
            for (size_t i = 1; i <= ido - 2; i += 2) {
                CH(i, k, j) = C1(i, k, j) - C1(i + 1, k, jc);
                CH(i, k, jc) = C1(i, k, j) + C1(i + 1, k, jc);
                CH(i + 1, k, j) = C1(i + 1, k, j) + C1(i, k, jc);
                CH(i + 1, k, jc) = C1(i + 1, k, j) - C1(i, k, jc);
            }
    for (size_t j = 1; j < ip; ++j) {
        size_t is = (j - 1) * (ido - 1);
        for (size_t k = 0; k < l1; ++k) {
            size_t idij = is;
            for (size_t i = 1; i <= ido - 2; i += 2) {
                float t1 = CH(i, k, j), t2 = CH(i + 1, k, j);
                CH(i, k, j) = wa[idij] * t1 - wa[idij + 1] * t2;
                CH(i + 1, k, j) = wa[idij] * t2 + wa[idij + 1] * t1;
                idij += 2;
            }
        }
    }
#undef CC
#undef CH
#undef C1
#undef C2
#undef CH2
}
#undef WA

/* ================================================================ cfftp<float> */
typedef struct { size_t fct; cf *tw, *tws; } cfct;
typedef struct { size_t length, nfct; cfct fct[MAXFACT]; cf *mem; } cfftp_t;
/* complex Bluestein pass (ducc0 cfftpblue), defined after fftblue below */
static void cblue_pass(size_t ido, size_t ip, size_t l1, cf *cc, cf *ch, const cf *wa, int fwd);

static int cfftp_init(cfftp_t *p, size_t length)
{
    p->length = length;
    p->nfct = 0;
    p->mem = NULL;
    if (length == 1) return 0;
    size_t len = length;
    while ((len & 7) == 0) { p->fct[p->nfct++].fct = 8; len >>= 3; }
    while ((len & 3) == 0) { p->fct[p->nfct++].fct = 4; len >>= 2; }
    if ((len & 1) == 0) {
        len >>= 1;
        p->fct[p->nfct++].fct = 2;
        size_t t = p->fct[0].fct; p->fct[0].fct = p->fct[p->nfct - 1].fct; p->fct[p->nfct - 1].fct = t;
    }
    for (size_t divisor = 3; divisor * divisor <= len; divisor += 2)
        while ((len % divisor) == 0) { p->fct[p->nfct++].fct = divisor; len /= divisor; }
    if (len > 1) p->fct[p->nfct++].fct = len;
    if (orc_ovr_cn) { size_t pr = 1; for (size_t i = 0; i < orc_ovr_cn; ++i) pr *= orc_ovr_c[i]; if (pr == length) { p->nfct = orc_ovr_cn; for (size_t i = 0; i < orc_ovr_cn; ++i) p->fct[i].fct = orc_ovr_c[i]; } }
    size_t twsz = 0, l1 = 1;
    for (size_t k = 0; k < p->nfct; ++k) {
        size_t ip = p->fct[k].fct, ido = length / (l1 * ip);
        twsz += (ip - 1) * (ido - 1);
        if (ip > 11) twsz += ip;
        l1 *= ip;
    }
    p->mem = (cf *)calloc(twsz + 1, sizeof(cf));
    sincos_t comp; sc_init(&comp, length);
    l1 = 1;
    size_t memofs = 0;
    for (size_t k = 0; k < p->nfct; ++k) {
        size_t ip = p->fct[k].fct, ido = length / (l1 * ip);
        p->fct[k].tw = p->mem + memofs;
        memofs += (ip - 1) * (ido - 1);
        for (size_t j = 1; j < ip; ++j)
            for (size_t i = 1; i < ido; ++i) p->fct[k].tw[(j - 1) * (ido - 1) + i - 1] = sc_get(&comp, j * l1 * i);
        p->fct[k].tws = NULL;
        if (ip > 11) {
            p->fct[k].tws = p->mem + memofs;
            memofs += ip;
            for (size_t j = 0; j < ip; ++j) p->fct[k].tws[j] = sc_get(&comp, j * l1 * ido);
        }
        l1 *= ip;
    }
    sc_free(&comp);
    return 0;
}
static void cfftp_free(cfftp_t *p) { free(p->mem); }

static inline cf c_add(cf a, cf b) { cf r = {a.r + b.r, a.i + b.i}; return r; }
static inline cf c_sub(cf a, cf b) { cf r = {a.r - b.r, a.i - b.i}; return r; }
/* special_mul<fwd>: fwd ? v1 * conj(v2) : v1 * v2 */
static inline cf smul(cf v1, cf v2, int fwd)
{
    cf r;
    if (fwd) { r.r = v1.r * v2.r + v1.i * v2.i; r.i = v1.i * v2.r - v1.r * v2.i; }
    else { r.r = v1.r * v2.r - v1.i * v2.i; r.i = v1.r * v2.i + v1.i * v2.r; }
    return r;
}
static inline cf rotx90(cf a, int fwd)
{
    cf r;
    if (fwd) { r.r = a.i; r.i = -a.r; } else { r.r = -a.i; r.i = a.r; }
    return r;
}
#define CPM(a, b, c, d) { a = c_add(c, d); b = c_sub(c, d); }
#define CWA(x, i) wa[(i) - 1 + (x) * (ido - 1)]
#define CCH(a, b, c) ch[(a) + ido * ((b) + l1 * (c))]

static void pass2(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
#define CCC(a, b, c) cc[(a) + ido * ((b) + 2 * (c))]
    for (size_t k = 0; k < l1; ++k) {
        CCH(0, k, 0) = c_add(CCC(0, 0, k), CCC(0, 1, k));
        CCH(0, k, 1) = c_sub(CCC(0, 0, k), CCC(0, 1, k));
        for (size_t i = 1; i < ido; ++i) {
            CCH(i, k, 0) = c_add(CCC(i, 0, k), CCC(i, 1, k));
            CCH(i, k, 1) = smul(c_sub(CCC(i, 0, k), CCC(i, 1, k)), CWA(0, i), fwd);
        }
    }
#undef CCC
}

static void pass3(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
    const float tw1r = -0.5f, tw1i = (fwd ? -1.f : 1.f) * (float)0.8660254037844386467637231707529362L;
#define CCC(a, b, c) cc[(a) + ido * ((b) + 3 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf t0 = CCC(i, 0, k), t1, t2;
            CPM(t1, t2, CCC(i, 1, k), CCC(i, 2, k))
            CCH(i, k, 0) = c_add(t0, t1);
            cf ca = {t0.r + t1.r * tw1r, t0.i + t1.i * tw1r};
            cf cb = {-t2.i * tw1i, t2.r * tw1i};
            if (i == 0) { CPM(CCH(0, k, 1), CCH(0, k, 2), ca, cb) }
            else {
                CCH(i, k, 1) = smul(c_add(ca, cb), CWA(0, i), fwd);
                CCH(i, k, 2) = smul(c_sub(ca, cb), CWA(1, i), fwd);
            }
        }
#undef CCC
}

static void pass4(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
#define CCC(a, b, c) cc[(a) + ido * ((b) + 4 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf t1, t2, t3, t4;
            CPM(t2, t1, CCC(i, 0, k), CCC(i, 2, k))
            CPM(t3, t4, CCC(i, 1, k), CCC(i, 3, k))
            t4 = rotx90(t4, fwd);
            if (i == 0) {
                CPM(CCH(0, k, 0), CCH(0, k, 2), t2, t3)
                CPM(CCH(0, k, 1), CCH(0, k, 3), t1, t4)
            } else {
                CCH(i, k, 0) = c_add(t2, t3);
                CCH(i, k, 1) = smul(c_add(t1, t4), CWA(0, i), fwd);
                CCH(i, k, 2) = smul(c_sub(t2, t3), CWA(1, i), fwd);
                CCH(i, k, 3) = smul(c_sub(t1, t4), CWA(2, i), fwd);
            }
        }
#undef CCC
}

static void pass5(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
    const float tw1r = (float)0.3090169943749474241022934171828191L,
                tw1i = (fwd ? -1.f : 1.f) * (float)0.9510565162951535721164393333793821L,
                tw2r = (float)-0.8090169943749474241022934171828191L,
                tw2i = (fwd ? -1.f : 1.f) * (float)0.5877852522924731291687059546390728L;
#define CCC(a, b, c) cc[(a) + ido * ((b) + 5 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf t0 = CCC(i, 0, k), t1, t2, t3, t4;
            CPM(t1, t4, CCC(i, 1, k), CCC(i, 4, k))
            CPM(t2, t3, CCC(i, 2, k), CCC(i, 3, k))
            CCH(i, k, 0).r = t0.r + t1.r + t2.r;
            CCH(i, k, 0).i = t0.i + t1.i + t2.i;
#define STEP5(u1, u2, twar, twbr, twai, twbi)                                   \
            {                                                                   \
                cf ca, cb;                                                      \
                ca.r = t0.r + twar * t1.r + twbr * t2.r;                        \
                ca.i = t0.i + twar * t1.i + twbr * t2.i;                        \
                cb.i = twai * t4.r twbi * t3.r;                                 \
                cb.r = -(twai * t4.i twbi * t3.i);                              \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }       \
                else {                                                          \
                    CCH(i, k, u1) = smul(c_add(ca, cb), CWA(u1 - 1, i), fwd);   \
                    CCH(i, k, u2) = smul(c_sub(ca, cb), CWA(u2 - 1, i), fwd);   \
                }                                                               \
            }
            STEP5(1, 4, tw1r, tw2r, +tw1i, +tw2i)
            STEP5(2, 3, tw2r, tw1r, +tw2i, -tw1i)
#undef STEP5
        }
#undef CCC
}

static void pass7(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
    const float tw1r = (float)0.6234898018587335305250048840042398L,
                tw1i = (fwd ? -1.f : 1.f) * (float)0.7818314824680298087084445266740578L,
                tw2r = (float)-0.2225209339563144042889025644967948L,
                tw2i = (fwd ? -1.f : 1.f) * (float)0.9749279121818236070181316829939312L,
                tw3r = (float)-0.9009688679024191262361023195074451L,
                tw3i = (fwd ? -1.f : 1.f) * (float)0.433883739117558120475768332848359L;
#define CCC(a, b, c) cc[(a) + ido * ((b) + 7 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf t1 = CCC(i, 0, k), t2, t3, t4, t5, t6, t7;
            CPM(t2, t7, CCC(i, 1, k), CCC(i, 6, k))
            CPM(t3, t6, CCC(i, 2, k), CCC(i, 5, k))
            CPM(t4, t5, CCC(i, 3, k), CCC(i, 4, k))
            CCH(i, k, 0).r = t1.r + t2.r + t3.r + t4.r;
            CCH(i, k, 0).i = t1.i + t2.i + t3.i + t4.i;
#define STEP7(u1, u2, x1, x2, x3, y1, y2, y3)                                   \
            {                                                                   \
                cf ca, cb;                                                      \
                ca.r = t1.r + x1 * t2.r + x2 * t3.r + x3 * t4.r;                \
                ca.i = t1.i + x1 * t2.i + x2 * t3.i + x3 * t4.i;                \
                cb.i = y1 * t7.r y2 * t6.r y3 * t5.r;                           \
                cb.r = -(y1 * t7.i y2 * t6.i y3 * t5.i);                        \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }       \
                else {                                                          \
                    CCH(i, k, u1) = smul(c_add(ca, cb), CWA(u1 - 1, i), fwd);   \
                    CCH(i, k, u2) = smul(c_sub(ca, cb), CWA(u2 - 1, i), fwd);   \
                }                                                               \
            }
            STEP7(1, 6, tw1r, tw2r, tw3r, +tw1i, +tw2i, +tw3i)
            STEP7(2, 5, tw2r, tw3r, tw1r, +tw2i, -tw3i, -tw1i)
            STEP7(3, 4, tw3r, tw1r, tw2r, +tw3i, -tw1i, +tw2i)
#undef STEP7
        }
#undef CCC
}

static inline cf rotx45(cf a, int fwd)
{
    const float hsqt2 = (float)0.707106781186547524400844362104849L;
    cf r;
    if (fwd) { float tmp_ = a.r; r.r = hsqt2 * (a.r + a.i); r.i = hsqt2 * (a.i - tmp_); }
    else { float tmp_ = a.r; r.r = hsqt2 * (a.r - a.i); r.i = hsqt2 * (a.i + tmp_); }
    return r;
}
static inline cf rotx135(cf a, int fwd)
{
    const float hsqt2 = (float)0.707106781186547524400844362104849L;
    cf r;
    if (fwd) { float tmp_ = a.r; r.r = hsqt2 * (a.i - a.r); r.i = hsqt2 * (-tmp_ - a.i); }
    else { float tmp_ = a.r; r.r = hsqt2 * (-a.r - a.i); r.i = hsqt2 * (tmp_ - a.i); }
    return r;
}

static void pass8(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
#define CCC(a, b, c) cc[(a) + ido * ((b) + 8 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf a0, a1, a2, a3, a4, a5, a6, a7;
            CPM(a1, a5, CCC(i, 1, k), CCC(i, 5, k))
            CPM(a3, a7, CCC(i, 3, k), CCC(i, 7, k))
            { cf s = c_add(a1, a3), d = c_sub(a1, a3); a1 = s; a3 = d; }
            a3 = rotx90(a3, fwd);
            a7 = rotx90(a7, fwd);
            { cf s = c_add(a5, a7), d = c_sub(a5, a7); a5 = s; a7 = d; }
            a5 = rotx45(a5, fwd);
            a7 = rotx135(a7, fwd);
            CPM(a0, a4, CCC(i, 0, k), CCC(i, 4, k))
            CPM(a2, a6, CCC(i, 2, k), CCC(i, 6, k))
            if (i == 0) {
                CPM(CCH(0, k, 0), CCH(0, k, 4), c_add(a0, a2), a1)
                CPM(CCH(0, k, 2), CCH(0, k, 6), c_sub(a0, a2), a3)
                a6 = rotx90(a6, fwd);
                CPM(CCH(0, k, 1), CCH(0, k, 5), c_add(a4, a6), a5)
                CPM(CCH(0, k, 3), CCH(0, k, 7), c_sub(a4, a6), a7)
            } else {
                { cf s = c_add(a0, a2), d = c_sub(a0, a2); a0 = s; a2 = d; }
                CCH(i, k, 0) = c_add(a0, a1);
                CCH(i, k, 4) = smul(c_sub(a0, a1), CWA(3, i), fwd);
                CCH(i, k, 2) = smul(c_add(a2, a3), CWA(1, i), fwd);
                CCH(i, k, 6) = smul(c_sub(a2, a3), CWA(5, i), fwd);
                a6 = rotx90(a6, fwd);
                { cf s = c_add(a4, a6), d = c_sub(a4, a6); a4 = s; a6 = d; }
                CCH(i, k, 1) = smul(c_add(a4, a5), CWA(0, i), fwd);
                CCH(i, k, 5) = smul(c_sub(a4, a5), CWA(4, i), fwd);
                CCH(i, k, 3) = smul(c_add(a6, a7), CWA(2, i), fwd);
                CCH(i, k, 7) = smul(c_sub(a6, a7), CWA(6, i), fwd);
            }
        }
#undef CCC
}

static void pass11(size_t ido, size_t l1, const cf *cc, cf *ch, const cf *wa, int fwd)
{
    const float tw1r = (float)0.8412535328311811688618116489193677L,
                tw1i = (fwd ? -1.f : 1.f) * (float)0.5406408174555975821076359543186917L,
                tw2r = (float)0.4154150130018864255292741492296232L,
                tw2i = (fwd ? -1.f : 1.f) * (float)0.9096319953545183714117153830790285L,
                tw3r = (float)-0.1423148382732851404437926686163697L,
                tw3i = (fwd ? -1.f : 1.f) * (float)0.9898214418809327323760920377767188L,
                tw4r = (float)-0.6548607339452850640569250724662936L,
                tw4i = (fwd ? -1.f : 1.f) * (float)0.7557495743542582837740358439723444L,
                tw5r = (float)-0.9594929736144973898903680570663277L,
                tw5i = (fwd ? -1.f : 1.f) * (float)0.2817325568414296977114179153466169L;
#define CCC(a, b, c) cc[(a) + ido * ((b) + 11 * (c))]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf t1 = CCC(i, 0, k), t2, t3, t4, t5, t6, t7, t8, t9, t10, t11;
            CPM(t2, t11, CCC(i, 1, k), CCC(i, 10, k))
            CPM(t3, t10, CCC(i, 2, k), CCC(i, 9, k))
            CPM(t4, t9, CCC(i, 3, k), CCC(i, 8, k))
            CPM(t5, t8, CCC(i, 4, k), CCC(i, 7, k))
            CPM(t6, t7, CCC(i, 5, k), CCC(i, 6, k))
            CCH(i, k, 0).r = t1.r + t2.r + t3.r + t4.r + t5.r + t6.r;
            CCH(i, k, 0).i = t1.i + t2.i + t3.i + t4.i + t5.i + t6.i;
#define STEP11(u1, u2, x1, x2, x3, x4, x5, y1, y2, y3, y4, y5)                              \
            {                                                                               \
                cf ca, cb;                                                                  \
                ca.r = t1.r + x1 * t2.r + x2 * t3.r + x3 * t4.r + x4 * t5.r + x5 * t6.r;    \
                ca.i = t1.i + x1 * t2.i + x2 * t3.i + x3 * t4.i + x4 * t5.i + x5 * t6.i;    \
                cb.i = y1 * t11.r y2 * t10.r y3 * t9.r y4 * t8.r y5 * t7.r;                 \
                cb.r = -(y1 * t11.i y2 * t10.i y3 * t9.i y4 * t8.i y5 * t7.i);              \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }                   \
                else {                                                                      \
                    CCH(i, k, u1) = smul(c_add(ca, cb), CWA(u1 - 1, i), fwd);               \
                    CCH(i, k, u2) = smul(c_sub(ca, cb), CWA(u2 - 1, i), fwd);               \
                }                                                                           \
            }
            STEP11(1, 10, tw1r, tw2r, tw3r, tw4r, tw5r, +tw1i, +tw2i, +tw3i, +tw4i, +tw5i)
            STEP11(2, 9, tw2r, tw4r, tw5r, tw3r, tw1r, +tw2i, +tw4i, -tw5i, -tw3i, -tw1i)
            STEP11(3, 8, tw3r, tw5r, tw2r, tw1r, tw4r, +tw3i, -tw5i, -tw2i, +tw1i, +tw4i)
            STEP11(4, 7, tw4r, tw3r, tw1r, tw5r, tw2r, +tw4i, -tw3i, +tw1i, +tw5i, -tw2i)
            STEP11(5, 6, tw5r, tw1r, tw4r, tw2r, tw3r, +tw5i, -tw1i, +tw4i, -tw2i, +tw3i)
#undef STEP11
        }
#undef CCC
}

/* generic complex radix (cfftp::passg / ducc0 cfftpg); the result ends in cc */
static void passg(size_t ido, size_t ip, size_t l1, cf *cc, cf *ch, const cf *wa, const cf *csarr, int fwd)
{
    const size_t cdim = ip, ipph = (ip + 1) / 2, idl1 = ido * l1;
#define GCH(a, b, c) ch[(a) + ido * ((b) + l1 * (c))]
#define GCC(a, b, c) cc[(a) + ido * ((b) + cdim * (c))]
#define GCX(a, b, c) cc[(a) + ido * ((b) + l1 * (c))]
#define GCX2(a, b) cc[(a) + idl1 * (b)]
#define GCH2(a, b) ch[(a) + idl1 * (b)]
    cf *wal = (cf *)malloc(sizeof(cf) * ip);
    wal[0].r = 1.f; wal[0].i = 0.f;
    for (size_t i = 1; i < ip; ++i) { wal[i].r = csarr[i].r; wal[i].i = fwd ? -csarr[i].i : csarr[i].i; }
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) GCH(i, k, 0) = GCC(i, 0, k);
    for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
        for (size_t k = 0; k < l1; ++k)
            for (size_t i = 0; i < ido; ++i) CPM(GCH(i, k, j), GCH(i, k, jc), GCC(i, j, k), GCC(i, jc, k))
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            cf tmp = GCH(i, k, 0);
            for (size_t j = 1; j < ipph; ++j) { tmp.r += GCH(i, k, j).r; tmp.i += GCH(i, k, j).i; }
            GCX(i, k, 0) = tmp;
        }
    for (size_t l = 1, lc = ip - 1; l < ipph; ++l, --lc) {
        for (size_t ik = 0; ik < idl1; ++ik) {
            GCX2(ik, l).r = GCH2(ik, 0).r + wal[l].r * GCH2(ik, 1).r + wal[2 * l].r * GCH2(ik, 2).r;
            GCX2(ik, l).i = GCH2(ik, 0).i + wal[l].r * GCH2(ik, 1).i + wal[2 * l].r * GCH2(ik, 2).i;
            GCX2(ik, lc).r = -(wal[l].i * GCH2(ik, ip - 1).i + wal[2 * l].i * GCH2(ik, ip - 2).i);
            GCX2(ik, lc).i = wal[l].i * GCH2(ik, ip - 1).r + wal[2 * l].i * GCH2(ik, ip - 2).r;
        }
        size_t iwal = 2 * l;
        size_t j = 3, jc = ip - 3;
        for (; j + 1 < ipph; j += 2, jc -= 2) {
            iwal += l; if (iwal > ip) iwal -= ip;
            cf xwal = wal[iwal];
            iwal += l; if (iwal > ip) iwal -= ip;
            cf xwal2 = wal[iwal];
            for (size_t ik = 0; ik < idl1; ++ik) {
                GCX2(ik, l).r += GCH2(ik, j).r * xwal.r + GCH2(ik, j + 1).r * xwal2.r;
                GCX2(ik, l).i += GCH2(ik, j).i * xwal.r + GCH2(ik, j + 1).i * xwal2.r;
                GCX2(ik, lc).r -= GCH2(ik, jc).i * xwal.i + GCH2(ik, jc - 1).i * xwal2.i;
                GCX2(ik, lc).i += GCH2(ik, jc).r * xwal.i + GCH2(ik, jc - 1).r * xwal2.i;
            }
        }
        for (; j < ipph; ++j, --jc) {
            iwal += l; if (iwal > ip) iwal -= ip;
            cf xwal = wal[iwal];
            for (size_t ik = 0; ik < idl1; ++ik) {
                GCX2(ik, l).r += GCH2(ik, j).r * xwal.r;
                GCX2(ik, l).i += GCH2(ik, j).i * xwal.r;
                GCX2(ik, lc).r -= GCH2(ik, jc).i * xwal.i;
                GCX2(ik, lc).i += GCH2(ik, jc).r * xwal.i;
            }
        }
    }
    if (ido == 1) {
        for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
            for (size_t ik = 0; ik < idl1; ++ik) {
                cf t1 = GCX2(ik, j), t2 = GCX2(ik, jc);
                CPM(GCX2(ik, j), GCX2(ik, jc), t1, t2)
            }
    } else {
        for (size_t j = 1, jc = ip - 1; j < ipph; ++j, --jc)
            for (size_t k = 0; k < l1; ++k) {
                cf t1 = GCX(0, k, j), t2 = GCX(0, k, jc);
                CPM(GCX(0, k, j), GCX(0, k, jc), t1, t2)
                for (size_t i = 1; i < ido; ++i) {
                    cf x1, x2;
                    CPM(x1, x2, GCX(i, k, j), GCX(i, k, jc))
                    size_t idij = (j - 1) * (ido - 1) + i - 1;
                    GCX(i, k, j) = smul(x1, wa[idij], fwd);
                    idij = (jc - 1) * (ido - 1) + i - 1;
                    GCX(i, k, jc) = smul(x2, wa[idij], fwd);
                }
            }
    }
    free(wal);
}

static void cfftp_exec(const cfftp_t *p, cf *c, float fct, int fwd)
{
    size_t length = p->length;
    if (length == 1) { c[0].r *= fct; c[0].i *= fct; return; }
    size_t l1 = 1;
    cf *ch = (cf *)malloc(sizeof(cf) * length);
    cf *p1 = c, *p2 = ch, *t;
    for (size_t k1 = 0; k1 < p->nfct; k1++) {
        size_t ip = p->fct[k1].fct;
        size_t l2 = ip * l1;
        size_t ido = length / l2;
        if (ip == 4) pass4(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 8) pass8(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 2) pass2(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 3) pass3(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 5) pass5(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 7) pass7(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip == 11) pass11(ido, l1, p1, p2, p->fct[k1].tw, fwd);
        else if (ip >= 110) { cblue_pass(ido, ip, l1, p1, p2, p->fct[k1].tw, fwd); if (l1 == 1) { t = p1; p1 = p2; p2 = t; } }
        else { passg(ido, ip, l1, p1, p2, p->fct[k1].tw, p->fct[k1].tws, fwd); t = p1; p1 = p2; p2 = t; }
        t = p1; p1 = p2; p2 = t;
        l1 = l2;
    }
    if (p1 != c) {
        if (fct != 1.f) for (size_t i = 0; i < length; ++i) { c[i].r = p1[i].r * fct; c[i].i = p1[i].i * fct; }
        else memcpy(c, p1, sizeof(cf) * length);
    } else if (fct != 1.f)
        for (size_t i = 0; i < length; ++i) { c[i].r *= fct; c[i].i *= fct; }
    free(ch);
}

/* ================================================================ fftblue<float> */
typedef struct { size_t n, n2; cfftp_t plan; cf *bk, *bkf; } fftblue_t;

static int fftblue_init(fftblue_t *b, size_t length)
{
    b->n = length;
    b->n2 = good_size_cmplx(length * 2 - 1);
    if (cfftp_init(&b->plan, b->n2)) return -1;
    size_t n = b->n, n2 = b->n2;
    b->bk = (cf *)malloc(sizeof(cf) * (n + n2 / 2 + 1));
    b->bkf = b->bk + n;
    sincos_t tmp; sc_init(&tmp, 2 * n);
    b->bk[0].r = 1.f; b->bk[0].i = 0.f;
    size_t coeff = 0;
    for (size_t m = 1; m < n; ++m) {
        coeff += 2 * m - 1;
        if (coeff >= 2 * n) coeff -= 2 * n;
        b->bk[m] = sc_get(&tmp, coeff);
    }
    sc_free(&tmp);
    cf *tbkf = (cf *)malloc(sizeof(cf) * n2);
    float xn2 = 1.f / (float)n2;
    tbkf[0].r = b->bk[0].r * xn2; tbkf[0].i = b->bk[0].i * xn2;
    for (size_t m = 1; m < n; ++m) {
        cf v = {b->bk[m].r * xn2, b->bk[m].i * xn2};
        tbkf[m] = tbkf[n2 - m] = v;
    }
    for (size_t m = n; m <= (n2 - n); ++m) { tbkf[m].r = 0.f; tbkf[m].i = 0.f; }
    cfftp_exec(&b->plan, tbkf, 1.f, 1);
    for (size_t i = 0; i < n2 / 2 + 1; ++i) b->bkf[i] = tbkf[i];
    free(tbkf);
    return 0;
}
static void fftblue_free(fftblue_t *b) { cfftp_free(&b->plan); free(b->bk); }

static void fftblue_fft(const fftblue_t *b, cf *c, float fct, int fwd)
{
    size_t n = b->n, n2 = b->n2;
    cf *akf = (cf *)malloc(sizeof(cf) * n2);
    for (size_t m = 0; m < n; ++m) akf[m] = smul(c[m], b->bk[m], fwd);
    cf zero = {akf[0].r * 0.f, akf[0].i * 0.f};
    for (size_t m = n; m < n2; ++m) akf[m] = zero;
    cfftp_exec(&b->plan, akf, 1.f, 1);
    akf[0] = smul(akf[0], b->bkf[0], !fwd);
    for (size_t m = 1; m < (n2 + 1) / 2; ++m) {
        akf[m] = smul(akf[m], b->bkf[m], !fwd);
        akf[n2 - m] = smul(akf[n2 - m], b->bkf[m], !fwd);
    }
    if ((n2 & 1) == 0) akf[n2 / 2] = smul(akf[n2 / 2], b->bkf[n2 / 2], !fwd);
    cfftp_exec(&b->plan, akf, 1.f, 0);
    for (size_t m = 0; m < n; ++m) {
        cf v = smul(akf[m], b->bk[m], fwd);
        c[m].r = v.r * fct; c[m].i = v.i * fct;
    }
    free(akf);
}

__attribute__((unused)) static void fftblue_exec_r(const fftblue_t *b, float *c, float fct, int fwd)
{
    size_t n = b->n;
    cf *tmp = (cf *)malloc(sizeof(cf) * n);
    if (fwd) {
        float zero = 0.f * c[0];
        for (size_t m = 0; m < n; ++m) { tmp[m].r = c[m]; tmp[m].i = zero; }
        fftblue_fft(b, tmp, fct, 1);
        c[0] = tmp[0].r;
        memcpy(c + 1, (float *)(tmp + 1), (n - 1) * sizeof(float));
    } else {
        tmp[0].r = c[0]; tmp[0].i = c[0] * 0.f;
        memcpy((float *)(tmp + 1), c + 1, (n - 1) * sizeof(float));
        if ((n & 1) == 0) tmp[n / 2].i = 0.f * c[0];
        for (size_t m = 1; 2 * m < n; ++m) { tmp[n - m].r = tmp[m].r; tmp[n - m].i = -tmp[m].i; }
        fftblue_fft(b, tmp, fct, 0);
        for (size_t m = 0; m < n; ++m) c[m] = tmp[m].r;
    }
    free(tmp);
}


/* ---- ducc0 cfftpblue<float> as a pass of a complex plan (prime radix >= 110): every (k, i) column is one Bluestein
 * transform of length ip; the inter-pass twiddle is merged into the final multiplication by b_k.  l1 > 1: result in ch;
 * l1 == 1: written back into cc. */
static void cblue_pass(size_t ido, size_t ip, size_t l1, cf *cc, cf *ch, const cf *wa, int fwd)
{
    fftblue_t b;
    if (fftblue_init(&b, ip)) return;
    size_t n2 = b.n2;
    cf *akf = (cf *)malloc(sizeof(cf) * n2);
#define BCC(a, b_, c) cc[(a) + ido * ((b_) + ip * (c))]
#define BCH(a, b_, c) ch[(a) + ido * ((b_) + l1 * (c))]
#define BWA(x, i) wa[(i) - 1 + (x) * (ido - 1)]
    for (size_t k = 0; k < l1; ++k)
        for (size_t i = 0; i < ido; ++i) {
            for (size_t m = 0; m < ip; ++m) akf[m] = smul(BCC(i, m, k), b.bk[m], fwd);
            cf zero = {akf[0].r * 0.f, akf[0].i * 0.f};
            for (size_t m = ip; m < n2; ++m) akf[m] = zero;
            cfftp_exec(&b.plan, akf, 1.f, 1);
            akf[0] = smul(akf[0], b.bkf[0], !fwd);
            for (size_t m = 1; m < (n2 + 1) / 2; ++m) {
                akf[m] = smul(akf[m], b.bkf[m], !fwd);
                akf[n2 - m] = smul(akf[n2 - m], b.bkf[m], !fwd);
            }
            if ((n2 & 1) == 0) akf[n2 / 2] = smul(akf[n2 / 2], b.bkf[n2 / 2], !fwd);
            cfftp_exec(&b.plan, akf, 1.f, 0);
            for (size_t m = 0; m < ip; ++m) {
                cf w = b.bk[m];
                if (i != 0 && m != 0) {
                    cf t = BWA(m - 1, i);
                    cf p = {b.bk[m].r * t.r - b.bk[m].i * t.i, b.bk[m].r * t.i + b.bk[m].i * t.r};
                    w = p;
                }
                cf v = smul(akf[m], w, fwd);
                if (l1 > 1) BCH(i, k, m) = v; else BCC(i, m, 0) = v;
            }
        }
#undef BCC
#undef BCH
#undef BWA
    free(akf);
    fftblue_free(&b);
}

/* ---- ducc0 rfftpblue<float>: one real-FFT pass of prime radix ip >= 135 evaluated with complex Bluestein transforms
 * (scipy >= 1.15 vendors ducc0, whose rfftpass::make_pass uses rfftpg for ip < 135 and rfftpblue above) */
#define WA(x, i) wa[(i) + (x) * (ido - 1)]
static void rblue_fwd(const fftblue_t *b, size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    size_t ip = b->n, ipph = (ip + 1) / 2;
    cf *cc2 = (cf *)malloc(sizeof(cf) * ip);
#define CC(a, b_, c) cc[(a) + ido * ((b_) + l1 * (c))]
#define CH(a, b_, c) ch[(a) + ido * ((b_) + ip * (c))]
    for (size_t k = 0; k < l1; ++k) {
        for (size_t m = 0; m < ip; ++m) { cc2[m].r = CC(0, k, m); cc2[m].i = 0.f; }
        fftblue_fft(b, cc2, 1.f, 1);
        CH(0, 0, k) = cc2[0].r;
        for (size_t m = 1; m <= ip / 2; ++m) {
            CH(ido - 1, 2 * m - 1, k) = cc2[m].r;
            CH(0, 2 * m, k) = cc2[m].i;
        }
    }
    if (ido != 1)
        for (size_t k = 0; k < l1; ++k)
            for (size_t i = 2, ic = ido - 2; i < ido; i += 2, ic -= 2) {
                cc2[0].r = CC(i - 1, k, 0); cc2[0].i = CC(i, k, 0);
                for (size_t m = 1; m < ipph; ++m) {
                    MULPM(cc2[m].r, cc2[m].i, WA(m - 1, i - 2), WA(m - 1, i - 1), CC(i - 1, k, m), CC(i, k, m))
                    MULPM(cc2[ip - m].r, cc2[ip - m].i, WA(ip - m - 1, i - 2), WA(ip - m - 1, i - 1), CC(i - 1, k, ip - m), CC(i, k, ip - m))
                }
                fftblue_fft(b, cc2, 1.f, 1);
                CH(i - 1, 0, k) = cc2[0].r;
                CH(i, 0, k) = cc2[0].i;
                for (size_t m = 1; m < ipph; ++m) {
                    CH(i - 1, 2 * m, k) = cc2[m].r;
                    CH(ic - 1, 2 * m - 1, k) = cc2[ip - m].r;
                    CH(i, 2 * m, k) = cc2[m].i;
                    CH(ic, 2 * m - 1, k) = -cc2[ip - m].i;
                }
            }
#undef CC
#undef CH
    free(cc2);
}
static void rblue_bwd(const fftblue_t *b, size_t ido, size_t l1, const float *cc, float *ch, const float *wa)
{
    size_t ip = b->n, ipph = (ip + 1) / 2;
    cf *cc2 = (cf *)malloc(sizeof(cf) * ip);
#define CC(a, b_, c) cc[(a) + ido * ((b_) + ip * (c))]
#define CH(a, b_, c) ch[(a) + ido * ((b_) + l1 * (c))]
    for (size_t k = 0; k < l1; ++k) {
        cc2[0].r = CC(0, 0, k); cc2[0].i = 0.f;
        for (size_t m = 1; m <= ip / 2; ++m) {
            cc2[m].r = CC(ido - 1, 2 * m - 1, k); cc2[m].i = CC(0, 2 * m, k);
            cc2[ip - m].r = CC(ido - 1, 2 * m - 1, k); cc2[ip - m].i = -CC(0, 2 * m, k);
        }
        fftblue_fft(b, cc2, 1.f, 0);
        for (size_t m = 0; m < ip; ++m) CH(0, k, m) = cc2[m].r;
    }
    if (ido != 1)
        for (size_t k = 0; k < l1; ++k)
            for (size_t i = 2, ic = ido - 2; i < ido; i += 2, ic -= 2) {
                cc2[0].r = CC(i - 1, 0, k); cc2[0].i = CC(i, 0, k);
                for (size_t m = 1; m < ipph; ++m) {
                    cc2[m].r = CC(i - 1, 2 * m, k); cc2[m].i = CC(i, 2 * m, k);
                    cc2[ip - m].r = CC(ic - 1, 2 * m - 1, k); cc2[ip - m].i = -CC(ic, 2 * m - 1, k);
                }
                fftblue_fft(b, cc2, 1.f, 0);
                CH(i - 1, k, 0) = cc2[0].r;
                CH(i, k, 0) = cc2[0].i;
                for (size_t m = 1; m < ip; ++m)
                    MULPM(CH(i, k, m), CH(i - 1, k, m), WA(m - 1, i - 2), WA(m - 1, i - 1), cc2[m].i, cc2[m].r)
            }
#undef CC
#undef CH
    free(cc2);
}
#undef WA

static void rfftp_exec(const rfftp_t *p, float *c, float fct, int r2hc)
{
    size_t length = p->length;
    if (length == 1) { c[0] *= fct; return; }
    size_t nf = p->nfct;
    float *ch = (float *)malloc(sizeof(float) * length);
    float *p1 = c, *p2 = ch, *t;
    if (r2hc) {
        for (size_t k1 = 0, l1 = length; k1 < nf; ++k1) {
            size_t k = nf - k1 - 1;
            size_t ip = p->fct[k].fct;
            size_t ido = length / l1;
            l1 /= ip;
            if (ip == 4) radf4(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 2) radf2(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 3) radf3(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 5) radf5(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip >= 135) { fftblue_t b; fftblue_init(&b, ip); rblue_fwd(&b, ido, l1, p1, p2, p->fct[k].tw); fftblue_free(&b); }
            else { radfg(ido, ip, l1, p1, p2, p->fct[k].tw, p->fct[k].tws); t = p1; p1 = p2; p2 = t; }
            t = p1; p1 = p2; p2 = t;
        }
    } else {
        for (size_t k = 0, l1 = 1; k < nf; k++) {
            size_t ip = p->fct[k].fct, ido = length / (ip * l1);
            if (ip == 4) radb4(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 2) radb2(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 3) radb3(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip == 5) radb5(ido, l1, p1, p2, p->fct[k].tw);
            else if (ip >= 135) { fftblue_t b; fftblue_init(&b, ip); rblue_bwd(&b, ido, l1, p1, p2, p->fct[k].tw); fftblue_free(&b); }
            else radbg(ido, ip, l1, p1, p2, p->fct[k].tw, p->fct[k].tws);
            t = p1; p1 = p2; p2 = t;
            l1 *= ip;
        }
    }
    if (p1 != c) {
        if (fct != 1.f) for (size_t i = 0; i < length; ++i) c[i] = fct * p1[i];
        else memcpy(c, p1, sizeof(float) * length);
    } else if (fct != 1.f)
        for (size_t i = 0; i < length; ++i) c[i] *= fct;
    free(ch);
}

/* ---- ducc0 rfftp_complexify<float>: even lengths > 1000 run as one complex transform of half the length */
static int complexify_exec(float *c, size_t N, float fct, int fwd)
{
    size_t h = N / 2;
    cfftp_t plan;
    if (cfftp_init(&plan, h)) return -2;
    sincos_t roots; sc_init(&roots, N);
    cf *res = (cf *)malloc(sizeof(cf) * h);
    if (fwd) {
        memcpy(res, c, sizeof(cf) * h);
        cfftp_exec(&plan, res, 1.f, 1);
        float *r = (float *)malloc(sizeof(float) * N);
        r[0] = res[0].r + res[0].i;
        for (size_t i = 1, xi = h - 1; i <= xi; ++i, --xi) {
            cf xe = {res[i].r + res[xi].r, res[i].i - res[xi].i};          /* res[i] + conj(res[xi]) */
            cf t = {res[i].i + res[xi].i, res[xi].r - res[i].r};
            cf w = sc_get(&roots, i);                                       /* conj applied in the product */
            cf xo = {t.r * w.r + t.i * w.i, t.i * w.r - t.r * w.i};
            r[2 * i - 1] = 0.5f * (xe.r + xo.r);
            r[2 * i] = 0.5f * (xe.i + xo.i);
            r[2 * xi - 1] = 0.5f * (xe.r - xo.r);
            r[2 * xi] = 0.5f * (xo.i - xe.i);
        }
        r[N - 1] = res[0].r - res[0].i;
        for (size_t i = 0; i < N; ++i) c[i] = fct != 1.f ? r[i] * fct : r[i];
        free(r);
    } else {
        res[0].r = c[0] + c[N - 1]; res[0].i = c[0] - c[N - 1];
        for (size_t i = 1, xi = h - 1; i <= xi; ++i, --xi) {
            cf t1 = {c[2 * i - 1], c[2 * i]};
            cf t2 = {c[2 * xi - 1], -c[2 * xi]};
            cf xe = {t1.r + t2.r, t1.i + t2.i};
            cf d = {t1.r - t2.r, t1.i - t2.i};
            cf w = sc_get(&roots, i);
            cf xo = {d.r * w.r - d.i * w.i, d.r * w.i + d.i * w.r};
            res[i].r = xe.r - xo.i; res[i].i = xe.i + xo.r;
            res[xi].r = xe.r + xo.i; res[xi].i = -xe.i + xo.r;
        }
        cfftp_exec(&plan, res, 1.f, 0);
        float *r = (float *)res;
        for (size_t i = 0; i < N; ++i) c[i] = fct != 1.f ? r[i] * fct : r[i];
    }
    free(res);
    sc_free(&roots);
    cfftp_free(&plan);
    return 0;
}

/* ================================================================ pocketfft_r<float> + r2r_fftpack */
/* How scipy (ducc0) evaluates the float32 r2r transform of length n (rows processed in its 4-wide SIMD batches):
 *   0  real passes (rfftp; Bluestein passes for prime factors >= 135): every odd or <= 1000 length, and even lengths
 *      > 1000 whose half length is 5-smooth;
 *   1  half-length complex transform (rfftp_complexify): even lengths > 1000 whose half length has a prime factor >= 7
 *      (generic complex radices below 110, a complex Bluestein pass for a prime factor >= 110);
 *  -1  not restated: two Bluestein factors (lengths beyond 36 000).
 * Pinned against the installed scipy for every even length in (1000, 3400) by tests/test_oracle.py.  Known gap: the rows
 * scipy processes outside its SIMD batches (the last rows % 4 rows of an array) take another route when 8 divides the
 * half length (class 1) or the length (5-smooth, class 0) and round differently; this file and the GPU reproduce the
 * SIMD rows, i.e. all but at most 3 rows of a sub-band. */
static int orc_allow_all = 0, orc_force_class = -2;
void orc_fft_force_class(int c) { orc_force_class = c; }
void orc_fft_allow_all(int v) { orc_allow_all = v; }
int orc_fft_class(size_t n)
{
    if (orc_force_class > -2) return orc_force_class;
    if (n < 1) return -1;
    if (n <= 1000 || (n & 1)) return 0;
    size_t h = n / 2, big = 1, n_blue = 0;
    for (size_t p = 2; p * p <= h; ++p)
        while (h % p == 0) { if (p > big) big = p; if (p >= 110) ++n_blue; h /= p; }
    if (h > 1) { if (h > big) big = h; if (h >= 110) ++n_blue; }
    if (big <= 5) return 0;
    return (n_blue <= 1 || orc_allow_all) ? 1 : -1;
}
int orc_fft_mirrored(size_t n) { return orc_fft_class(n) >= 0; }

/* scipy.fftpack.rfft (forward != 0) / irfft (forward == 0, scaled by 1/n) on `rows` contiguous rows of length n.
 * Returns 0, or 1 when the length class is not restated (data untouched). */
int orc_fftpack_r2r_f32(float *data, size_t rows, size_t n, int forward)
{
    if (n == 0) return -1;
    const int cls = orc_fft_class(n);
    if (cls < 0) return 1;
    float fct = forward ? 1.f : (float)(1.0L / (long double)n);
    if (cls == 1) {
        for (size_t r = 0; r < rows; ++r) { int rc = complexify_exec(data + r * n, n, fct, forward); if (rc) return rc; }
        return 0;
    }
    rfftp_t p;
    rfftp_init(&p, n);
    for (size_t r = 0; r < rows; ++r) rfftp_exec(&p, data + r * n, fct, forward);
    rfftp_free(&p);
    return 0;
}

/* scipy.fft.fft / ifft(norm="forward"-less: unscaled) on complex64 rows, for pinning cfftp alone (11-smooth n only) */
int orc_cfft_c64(float *data, size_t rows, size_t n, int forward)
{
    cfftp_t p;
    if (cfftp_init(&p, n)) return -2;
    for (size_t r = 0; r < rows; ++r) cfftp_exec(&p, (cf *)data + r * n, 1.f, forward);
    cfftp_free(&p);
    return 0;
}
