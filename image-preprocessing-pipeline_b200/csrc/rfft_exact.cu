// Rounding-exact batched real FFT -> notch -> inverse real FFT in shared memory.
//
// Replaces np_filter_coefficient (pystripe/core.py:749-754): scipy.fftpack.rfft, multiplication by np_notch indexed by
// packed position, scipy.fftpack.irfft.  scipy (>= 1.15: scipy/fft/_duccfft, before: pocketfft) evaluates the float32
// transform as a sequence of real radix passes (4, 2, 3, 5, generic odd radix < 135) and, for prime factors >= 135, a
// Bluestein pass built on a complex FFT of 11-smooth length (complex radices 2,3,4,5,7,8,11).  To reproduce the
// reference's uint16 output bit for bit, the kernels below execute the SAME passes with the SAME float32 operation
// order (no FMA: this file is compiled with -fmad=false) and the same twiddle tables (double-precision two-table
// product rounded to float).  What is parallel here is what is independent there: rows of the sub-band and the
// (k, i) butterflies inside a pass.
//
// Layout: a CTA holds G sequences ("rows") in shared memory, element e of row r at (e * GP + r), GP = G + 1, so
//   * the global gather/scatter (consecutive e) and the passes (consecutive r) are both bank-conflict free,
//   * a warp works on 32/G butterflies x G rows: index arithmetic and twiddle loads are (nearly) warp-uniform.
// Even lengths > 1000 whose half length has a prime factor >= 7 run, as in ducc0, as a half-length complex transform
// (k_notch_cplx below: complex radices 2,3,4,5,7,8,11, generic, Bluestein).  k_notch (fft.cu) remains for exact=0.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b2s_internal.h"

namespace {

constexpr int kXT = 512;   // threads per CTA (compile-time: 448 / 384 / 320 / 256 measured 1 - 10 % slower, profiles/r02_notch_experiments.md)

struct Ctx { int r, item0, istride, GP; };
// one pass of the real transform.  kind: 2,3,4,5 = radix, 6 = generic, 7 = Bluestein; m_*: fdiv magics for l1, ni = (ido-1)/2,
// l1*ni, ido, ido*l1
struct XPass { int kind, ip, l1, ido, tw, cs; unsigned m_l1, m_ni, m_l1ni, m_ido, m_idl1; int lb; };
#define IDX(e) ((e) * c.GP + c.r)
#define FOR_ITEMS(it, count) for (int it = c.item0, cnt__ = (count); it < cnt__; it += c.istride)
#define PM(a, b, cc_, d) { a = (cc_) + (d); b = (cc_) - (d); }
#define MULPM(a, b, cc_, d, e, f) { a = (cc_) * (e) + (d) * (f); b = (cc_) * (f) - (d) * (e); }
#define WA(x, i) __ldg(wa + (i) + (x) * (ido - 1))

// ================================================================ real passes (mirror of rfftp radf*/radb*)
__device__ void radf2(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + l1 * (k_)))]
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + 2 * (k_)))]
    FOR_ITEMS(k, l1) {
        PM(CH(0, 0, k), CH(ido - 1, 1, k), CC(0, k, 0), CC(0, k, 1))
        if ((ido & 1) == 0) {
            CH(0, 1, k) = -CC(ido - 1, k, 1);
            CH(ido - 1, 0, k) = CC(ido - 1, k, 0);
        }
    }
    if (ido <= 2) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
        float tr2, ti2;
        MULPM(tr2, ti2, WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
        PM(CH(i - 1, 0, k), CH(ic - 1, 1, k), CC(i - 1, k, 0), tr2)
        PM(CH(i, 0, k), CH(ic, 1, k), ti2, CC(i, k, 0))
    }
#undef CH
}

__device__ void radf3(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float taur = -0.5f, taui = 0.8660254037844386467637231707529362f;
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + 3 * (k_)))]
    FOR_ITEMS(k, l1) {
        const float cr2 = CC(0, k, 1) + CC(0, k, 2);
        CH(0, 0, k) = CC(0, k, 0) + cr2;
        CH(0, 2, k) = taui * (CC(0, k, 2) - CC(0, k, 1));
        CH(ido - 1, 1, k) = CC(0, k, 0) + taur * cr2;
    }
    if (ido == 1) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
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

__device__ void radf4(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float hsqt2 = 0.707106781186547524400844362104849f;
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + 4 * (k_)))]
    FOR_ITEMS(k, l1) {
        float tr1, tr2;
        PM(tr1, CH(0, 2, k), CC(0, k, 3), CC(0, k, 1))
        PM(tr2, CH(ido - 1, 1, k), CC(0, k, 0), CC(0, k, 2))
        PM(CH(0, 0, k), CH(ido - 1, 3, k), tr2, tr1)
        if ((ido & 1) == 0) {
            const float ti1 = -hsqt2 * (CC(ido - 1, k, 1) + CC(ido - 1, k, 3));
            const float tr1b = hsqt2 * (CC(ido - 1, k, 1) - CC(ido - 1, k, 3));
            PM(CH(ido - 1, 0, k), CH(ido - 1, 2, k), CC(ido - 1, k, 0), tr1b)
            PM(CH(0, 3, k), CH(0, 1, k), ti1, CC(ido - 1, k, 2))
        }
    }
    if (ido <= 2) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
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

__device__ void radf5(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float tr11 = 0.3090169943749474241022934171828191f, ti11 = 0.9510565162951535721164393333793821f,
                tr12 = -0.8090169943749474241022934171828191f, ti12 = 0.5877852522924731291687059546390728f;
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + 5 * (k_)))]
    FOR_ITEMS(k, l1) {
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
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
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
        const float tr2 = CC(i - 1, k, 0) + tr11 * cr2 + tr12 * cr3;
        const float ti2 = CC(i, k, 0) + tr11 * ci2 + tr12 * ci3;
        const float tr3 = CC(i - 1, k, 0) + tr12 * cr2 + tr11 * cr3;
        const float ti3 = CC(i, k, 0) + tr12 * ci2 + tr11 * ci3;
        const float tr5 = cr5 * ti11 + cr4 * ti12;
        const float ti5 = ci5 * ti11 + ci4 * ti12;
        const float tr4 = cr5 * ti12 - cr4 * ti11;
        const float ti4 = ci5 * ti12 - ci4 * ti11;
        PM(CH(i - 1, 2, k), CH(ic - 1, 1, k), tr2, tr5)
        PM(CH(i, 2, k), CH(ic, 1, k), ti5, ti2)
        PM(CH(i - 1, 4, k), CH(ic - 1, 3, k), tr3, tr4)
        PM(CH(i, 4, k), CH(ic, 3, k), ti4, ti3)
    }
#undef CH
#undef CC
}

// The O(ip^2) part of a generic-radix pass: for every l in [1, ipph) and every ik
//   DST(ik, l)    = SRC(ik,0) + sum_j cos(2 pi l j / ip) SRC(ik, j)        (pocketfft's 4 / 2 / 1 grouping of j >= 3)
//   DST(ik, ip-l) =             sum_j sin(2 pi l j / ip) SRC(ik, ip - j)
// One thread owns LB consecutive l for one ik, so every SRC value it loads from shared memory feeds LB output pairs.
// The (cos, sin) pairs come from a per-pass table T[l-1][j-1] staged in shared memory (warp-uniform 128-bit loads) that
// holds csarr[2*((l*j) mod ip)] — the values pocketfft walks with its iang counter.  The cos-sum and the sin-sum of one l
// advance in lock step, so they share packed f32x2 instructions: products as fma(t, v, -0.0) (an exactly rounded
// product that ptxas cannot re-contract with the following add; -0.0 arrives as a kernel argument), sums as FADD2 in
// the reference's association order.
// exact integer division by a per-pass constant: q = umulhi(it, magic), magic = floor(2^32 / d) + 1 (0 encodes d == 1);
// valid while it * d < 2^32, which the plan builder checks
__device__ __forceinline__ int fdiv(int it, unsigned magic) { return magic ? (int)__umulhi((unsigned)it, magic) : it; }

constexpr int kLBMax = 8;   // outputs (l values) per thread in the O(radix^2) phase: 4 .. 8, chosen per pass (XPass::lb) so that the
                           // equally heavy items fill whole rounds of the CTA's item slots
constexpr int kRow = 68;   // float2 entries per table row (radix < 135 => at most 66 columns); compile-time so that the LB rows
                          // a thread reads sit at immediate offsets from one running pointer
__device__ __forceinline__ float2 mul2x(float2 t, float2 v, float2 nz) { return __ffma2_rn(t, v, nz); }
// MODE 0: store the pair (cos sum, sin sum) of output l at (ik, l) and (ik, ip - l) of dst.
// MODE 1 (radfg): the sums of one item are all that the rest of the forward pass needs — its shuffle into the r2hc layout
//   and, for i >= 1, the sum / difference of the pair — so they are stored straight at their final positions
//   OUT(a, b, k) = dst[a + ido * (b + ip * k)]:  i == 0: OUT(ido-1, 2l-1, k) = X, OUT(0, 2l, k) = Y;
//   i odd: OUT(i, 2l, k) = X + Y, OUT(ido-i-2, 2l-1, k) = X - Y;  i even: OUT(i, 2l, k) = X + Y, OUT(ido-i, 2l-1, k) = Y - X
//   (the same single add / subtract of the same two floats as pocketfft's last loop of radfg).
// MODE 2 (radbg with ido == 1): every item is an i == 0 item: dst(ik, l) = X - Y, dst(ik, ip - l) = X + Y (pocketfft's PM).
// INMODE 0: the inputs are (src(ik, j), src(ik, ip - j)).
// INMODE 1 (radfg with ido == 1): the PM that precedes the O(ip^2) phase is applied on the way in: (b + a, b - a) of the
//   same two elements a = src(ik, j), b = src(ik, ip - j).
// INMODE 2 (radbg with ido == 1): src is still in the r2hc layout: (2 * src[2j - 1], 2 * src[2j]) of block ik, x0 = src[0].
template <int INMODE>
__device__ __forceinline__ float2 gin(float a, float b)
{
    return INMODE == 0 ? make_float2(a, b) : (INMODE == 1 ? make_float2(b + a, b - a) : make_float2(2.f * a, 2.f * b));
}
template <int LB, int MODE, int INMODE = 0>
__device__ __forceinline__ void generic_block(const Ctx &c, const float *src, float *dst, const float2 *sgt, int ip,
                                              int ipph, int idl1, int lb, int ik, float2 nz, int ido = 1, unsigned m_ido = 0)
{
    const int l0 = 1 + LB * lb, nl = min(LB, ipph - l0);
    const int st = idl1 * c.GP;
    // pf[0] / pb[0]: the two inputs of index j = 1; they advance by sf / sb per j
    const int sf = INMODE == 2 ? 2 * c.GP : st, sb = INMODE == 2 ? 2 * c.GP : -st;
    const float *p0 = INMODE == 2 ? src + (ip * ik) * c.GP + c.r : src + ik * c.GP + c.r;
    const float *pf = INMODE == 2 ? p0 + c.GP : p0 + st;
    const float *pb = INMODE == 2 ? p0 + 2 * c.GP : p0 + (ip - 1) * st;
    float2 A[LB];
    // rows l0 .. l0+LB-1 of the table (rows past ipph-1 are zero padding: their results are not stored)
    const float4 *tb = reinterpret_cast<const float4 *>(sgt + (l0 - 1) * kRow);
    constexpr int RQ = kRow / 2;   // float4 per row
    {
        const float2 x0 = make_float2(p0[0], 0.f);
        const float2 v1 = gin<INMODE>(pf[0], pb[0]), v2 = gin<INMODE>(pf[sf], pb[sb]);
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const float4 t = tb[q * RQ];
            A[q] = __fadd2_rn(__fadd2_rn(x0, mul2x(make_float2(t.x, t.y), v1, nz)), mul2x(make_float2(t.z, t.w), v2, nz));
        }
    }
    pf += 2 * sf;
    pb += 2 * sb;
    tb += 1;
    int j = 3;
    for (; j + 3 < ipph; j += 4) {
        const float2 v0 = gin<INMODE>(pf[0], pb[0]), v1 = gin<INMODE>(pf[sf], pb[sb]);
        const float2 v2 = gin<INMODE>(pf[2 * sf], pb[2 * sb]), v3 = gin<INMODE>(pf[3 * sf], pb[3 * sb]);
        pf += 4 * sf;
        pb += 4 * sb;
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const float4 t = tb[q * RQ], u = tb[q * RQ + 1];
            float2 sacc = __fadd2_rn(mul2x(make_float2(t.x, t.y), v0, nz), mul2x(make_float2(t.z, t.w), v1, nz));
            sacc = __fadd2_rn(sacc, mul2x(make_float2(u.x, u.y), v2, nz));
            sacc = __fadd2_rn(sacc, mul2x(make_float2(u.z, u.w), v3, nz));
            A[q] = __fadd2_rn(A[q], sacc);
        }
        tb += 2;
    }
    for (; j + 1 < ipph; j += 2) {
        const float2 v0 = gin<INMODE>(pf[0], pb[0]), v1 = gin<INMODE>(pf[sf], pb[sb]);
        pf += 2 * sf;
        pb += 2 * sb;
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const float4 t = tb[q * RQ];
            A[q] = __fadd2_rn(A[q], __fadd2_rn(mul2x(make_float2(t.x, t.y), v0, nz), mul2x(make_float2(t.z, t.w), v1, nz)));
        }
        tb += 1;
    }
    if (j < ipph) {
        const float2 v0 = gin<INMODE>(pf[0], pb[0]);
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const float2 t = reinterpret_cast<const float2 *>(tb + q * RQ)[0];
            A[q] = __fadd2_rn(A[q], mul2x(t, v0, nz));
        }
    }
    if (MODE == 0) {
        float *pd = dst + ik * c.GP + c.r;
#pragma unroll
        for (int q = 0; q < LB; ++q)
            if (q < nl) { pd[(l0 + q) * st] = A[q].x; pd[(ip - l0 - q) * st] = A[q].y; }
    } else if (MODE == 2) {
        float *pd = dst + ik * c.GP + c.r;
#pragma unroll
        for (int q = 0; q < LB; ++q)
            if (q < nl) { pd[(l0 + q) * st] = A[q].x - A[q].y; pd[(ip - l0 - q) * st] = A[q].x + A[q].y; }
    } else {
        const int k = fdiv(ik, m_ido), i = ik - k * ido;
        // element (a, b) of block k: dst[(a + ido * (b + ip * k)) * GP + r]
        float *blk = dst + (ido * ip * k) * c.GP + c.r;
        const int rs = ido * c.GP;                       // one step of b
        int a1, a2;                                      // first index of the 2l row and of the 2l-1 row
        if (i == 0) { a1 = 0; a2 = ido - 1; }
        else if (i & 1) { a1 = i; a2 = ido - i - 2; }
        else { a1 = i; a2 = ido - i; }
        float *p1 = blk + a1 * c.GP, *p2 = blk + a2 * c.GP;
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            if (q < nl) {
                const int l = l0 + q;
                const float X = A[q].x, Y = A[q].y;
                float hi, lo;                            // -> row 2l, row 2l-1
                if (i == 0) { hi = Y; lo = X; }
                else if (i & 1) { hi = X + Y; lo = X - Y; }
                else { hi = X + Y; lo = Y - X; }
                p1[(2 * l) * rs] = hi;
                p2[(2 * l - 1) * rs] = lo;
            }
        }
    }
}

template <int MODE, int INMODE = 0>
__device__ __forceinline__ void generic_block_lb(int LBsel, const Ctx &c, const float *src, float *dst, const float2 *sgt, int ip,
                                                 int ipph, int idl1, int lb, int ik, float2 nz, int ido = 1, unsigned m_ido = 0)
{
    switch (LBsel) {
    case 4: generic_block<4, MODE, INMODE>(c, src, dst, sgt, ip, ipph, idl1, lb, ik, nz, ido, m_ido); break;
    case 5: generic_block<5, MODE, INMODE>(c, src, dst, sgt, ip, ipph, idl1, lb, ik, nz, ido, m_ido); break;
    case 6: generic_block<6, MODE, INMODE>(c, src, dst, sgt, ip, ipph, idl1, lb, ik, nz, ido, m_ido); break;
    case 7: generic_block<7, MODE, INMODE>(c, src, dst, sgt, ip, ipph, idl1, lb, ik, nz, ido, m_ido); break;
    default: generic_block<8, MODE, INMODE>(c, src, dst, sgt, ip, ipph, idl1, lb, ik, nz, ido, m_ido); break;
    }
}


// generic odd radix, forward; the result ends in ch.  `cs` is a shared-memory copy of csarr (2*ip floats).
__device__ void radfg(const Ctx &c, const XPass &P, float *cc, float *ch, const float *wa, const float2 *gt, float2 *sgt, float2 nz)
{
    const int ido = P.ido, ip = P.ip, l1 = P.l1;
    const int cdim = ip, ipph = (ip + 1) / 2, idl1 = ido * l1;
    const int LB = P.lb;
    const int nlb = (ipph - 1 + LB - 1) / LB;
    for (int i = threadIdx.x; i < nlb * LB * (kRow / 2); i += kXT)   // visible after the next barrier
        reinterpret_cast<float4 *>(sgt)[i] = __ldg(reinterpret_cast<const float4 *>(gt) + i);
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + cdim * (k_)))]
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + l1 * (k_)))]
#define C1(a, b, k_) cc[IDX((a) + ido * ((b) + l1 * (k_)))]
#define C2(a, b) cc[IDX((a) + idl1 * (b))]
#define CH2(a, b) ch[IDX((a) + idl1 * (b))]
    if (ido == 1) {
        // one phase: the PM of the (i == 0) elements rides on the loads (INMODE 1), the r2hc shuffle on the stores (MODE 1)
        __syncthreads();   // sgt
        FOR_ITEMS(it, (nlb + 1) * idl1) {
            const int lb = fdiv(it, P.m_idl1), ik = it - lb * idl1;
            if (lb == nlb) {
                float s = C2(ik, 0);
                for (int j = 1; j < ipph; ++j) s += C2(ik, ip - j) + C2(ik, j);      // C1(0,k,j) after the PM: t2 + t1
                ch[IDX(cdim * ik)] = s;
            } else generic_block_lb<1, 1>(LB, c, cc, ch, sgt, ip, ipph, idl1, lb, ik, nz, ido, P.m_ido);
        }
        return;
    }
    if (ido > 1) {
        const int ni = (ido - 1) >> 1;
        FOR_ITEMS(it, (ipph - 1) * l1 * ni) {
            const int jj = fdiv(it, P.m_l1ni), rem = it - jj * (l1 * ni);
            const int k = fdiv(rem, P.m_ni), i = 1 + 2 * (rem - k * ni);
            const int j = jj + 1, jc = ip - j;
            const int idij = (j - 1) * (ido - 1) + (i - 1), idij2 = (jc - 1) * (ido - 1) + (i - 1);
            const float t1 = C1(i, k, j), t2 = C1(i + 1, k, j), t3 = C1(i, k, jc), t4 = C1(i + 1, k, jc);
            const float w0 = __ldg(wa + idij), w1 = __ldg(wa + idij + 1), v0 = __ldg(wa + idij2), v1 = __ldg(wa + idij2 + 1);
            const float x1 = w0 * t1 + w1 * t2, x2 = w0 * t2 - w1 * t1, x3 = v0 * t3 + v1 * t4, x4 = v0 * t4 - v1 * t3;
            PM(C1(i, k, j), C1(i + 1, k, jc), x3, x1)
            PM(C1(i + 1, k, j), C1(i, k, jc), x2, x4)
        }
    }
    FOR_ITEMS(it, (ipph - 1) * l1) {
        const int jj = fdiv(it, P.m_l1), k = it - jj * l1, j = jj + 1, jc = ip - j;
        const float t1 = C1(0, k, j), t2 = C1(0, k, jc);
        PM(C1(0, k, j), C1(0, k, jc), t2, t1)
    }
    __syncthreads();
    // O(ip^2) phase; every item stores its results at their final r2hc positions in ch (see generic_block, MODE 1), the
    // l == 0 item (the plain sum) at OUT(i, 0, k): the result of the pass is in ch
    FOR_ITEMS(it, (nlb + 1) * idl1) {
        const int lb = fdiv(it, P.m_idl1), ik = it - lb * idl1;
        if (lb == nlb) {
            float s = C2(ik, 0);
            for (int j = 1; j < ipph; ++j) s += C2(ik, j);
            const int k = fdiv(ik, P.m_ido), i = ik - k * ido;
            ch[IDX(i + ido * (cdim * k))] = s;
        } else generic_block_lb<1>(LB, c, cc, ch, sgt, ip, ipph, idl1, lb, ik, nz, ido, P.m_ido);
    }
#undef CC
#undef CH
}

// ---- backward
__device__ void radb2(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + 2 * (k_)))]
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + l1 * (k_)))]
    FOR_ITEMS(k, l1) {
        PM(CH(0, k, 0), CH(0, k, 1), CC(0, 0, k), CC(ido - 1, 1, k))
        if ((ido & 1) == 0) {
            CH(ido - 1, k, 0) = 2.f * CC(ido - 1, 0, k);
            CH(ido - 1, k, 1) = -2.f * CC(0, 1, k);
        }
    }
    if (ido <= 2) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
        float ti2, tr2;
        PM(CH(i - 1, k, 0), tr2, CC(i - 1, 0, k), CC(ic - 1, 1, k))
        PM(ti2, CH(i, k, 0), CC(i, 0, k), CC(ic, 1, k))
        MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), ti2, tr2)
    }
#undef CC
}

__device__ void radb3(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float taur = -0.5f, taui = 0.8660254037844386467637231707529362f;
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + 3 * (k_)))]
    FOR_ITEMS(k, l1) {
        const float tr2 = 2.f * CC(ido - 1, 1, k);
        const float cr2 = CC(0, 0, k) + taur * tr2;
        CH(0, k, 0) = CC(0, 0, k) + tr2;
        const float ci3 = 2.f * taui * CC(0, 2, k);
        PM(CH(0, k, 2), CH(0, k, 1), cr2, ci3)
    }
    if (ido == 1) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
        const float tr2 = CC(i - 1, 2, k) + CC(ic - 1, 1, k);
        const float ti2 = CC(i, 2, k) - CC(ic, 1, k);
        const float cr2 = CC(i - 1, 0, k) + taur * tr2;
        const float ci2 = CC(i, 0, k) + taur * ti2;
        CH(i - 1, k, 0) = CC(i - 1, 0, k) + tr2;
        CH(i, k, 0) = CC(i, 0, k) + ti2;
        const float cr3 = taui * (CC(i - 1, 2, k) - CC(ic - 1, 1, k));
        const float ci3 = taui * (CC(i, 2, k) + CC(ic, 1, k));
        float di2, di3, dr2, dr3;
        PM(dr3, dr2, cr2, ci3)
        PM(di2, di3, ci2, cr3)
        MULPM(CH(i, k, 1), CH(i - 1, k, 1), WA(0, i - 2), WA(0, i - 1), di2, dr2)
        MULPM(CH(i, k, 2), CH(i - 1, k, 2), WA(1, i - 2), WA(1, i - 1), di3, dr3)
    }
#undef CC
}

__device__ void radb4(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float sqrt2 = 1.414213562373095048801688724209698f;
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + 4 * (k_)))]
    FOR_ITEMS(k, l1) {
        float tr1, tr2;
        PM(tr2, tr1, CC(0, 0, k), CC(ido - 1, 3, k))
        const float tr3 = 2.f * CC(ido - 1, 1, k);
        const float tr4 = 2.f * CC(0, 2, k);
        PM(CH(0, k, 0), CH(0, k, 2), tr2, tr3)
        PM(CH(0, k, 3), CH(0, k, 1), tr1, tr4)
        if ((ido & 1) == 0) {
            float ti1, ti2, ur1, ur2;
            PM(ti1, ti2, CC(0, 3, k), CC(0, 1, k))
            PM(ur2, ur1, CC(ido - 1, 0, k), CC(ido - 1, 2, k))
            CH(ido - 1, k, 0) = ur2 + ur2;
            CH(ido - 1, k, 1) = sqrt2 * (ur1 - ti1);
            CH(ido - 1, k, 2) = ti2 + ti2;
            CH(ido - 1, k, 3) = -sqrt2 * (ur1 + ti1);
        }
    }
    if (ido <= 2) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
        float ci2, ci3, ci4, cr2, cr3, cr4, ti1, ti2, ti3, ti4, tr1, tr2, tr3, tr4;
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

__device__ void radb5(const Ctx &c, int ido, int l1, const float *cc, float *ch, const float *wa)
{
    const float tr11 = 0.3090169943749474241022934171828191f, ti11 = 0.9510565162951535721164393333793821f,
                tr12 = -0.8090169943749474241022934171828191f, ti12 = 0.5877852522924731291687059546390728f;
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + 5 * (k_)))]
    FOR_ITEMS(k, l1) {
        const float ti5 = CC(0, 2, k) + CC(0, 2, k);
        const float ti4 = CC(0, 4, k) + CC(0, 4, k);
        const float tr2 = CC(ido - 1, 1, k) + CC(ido - 1, 1, k);
        const float tr3 = CC(ido - 1, 3, k) + CC(ido - 1, 3, k);
        CH(0, k, 0) = CC(0, 0, k) + tr2 + tr3;
        const float cr2 = CC(0, 0, k) + tr11 * tr2 + tr12 * tr3;
        const float cr3 = CC(0, 0, k) + tr12 * tr2 + tr11 * tr3;
        float ci4, ci5;
        MULPM(ci5, ci4, ti5, ti4, ti11, ti12)
        PM(CH(0, k, 4), CH(0, k, 1), cr2, ci5)
        PM(CH(0, k, 3), CH(0, k, 2), cr3, ci4)
    }
    if (ido == 1) return;
    const int ni = (ido - 1) >> 1;
    FOR_ITEMS(it, l1 * ni) {
        const int k = it / ni, i = 2 + 2 * (it - k * ni), ic = ido - i;
        float tr2, tr3, tr4, tr5, ti2, ti3, ti4, ti5;
        PM(tr2, tr5, CC(i - 1, 2, k), CC(ic - 1, 1, k))
        PM(ti5, ti2, CC(i, 2, k), CC(ic, 1, k))
        PM(tr3, tr4, CC(i - 1, 4, k), CC(ic - 1, 3, k))
        PM(ti4, ti3, CC(i, 4, k), CC(ic, 3, k))
        CH(i - 1, k, 0) = CC(i - 1, 0, k) + tr2 + tr3;
        CH(i, k, 0) = CC(i, 0, k) + ti2 + ti3;
        const float cr2 = CC(i - 1, 0, k) + tr11 * tr2 + tr12 * tr3;
        const float ci2 = CC(i, 0, k) + tr11 * ti2 + tr12 * ti3;
        const float cr3 = CC(i - 1, 0, k) + tr12 * tr2 + tr11 * tr3;
        const float ci3 = CC(i, 0, k) + tr12 * ti2 + tr11 * ti3;
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

// generic odd radix, backward; returns true when the result ends in ch
__device__ bool radbg(const Ctx &c, const XPass &P, float *cc, float *ch, const float *wa, const float2 *gt, float2 *sgt, float2 nz)
{
    const int ido = P.ido, ip = P.ip, l1 = P.l1;
    const int cdim = ip, ipph = (ip + 1) / 2, idl1 = ido * l1;
    const int LB = P.lb;
    const int nlb = (ipph - 1 + LB - 1) / LB;
    for (int i = threadIdx.x; i < nlb * LB * (kRow / 2); i += kXT)   // visible after the next barrier
        reinterpret_cast<float4 *>(sgt)[i] = __ldg(reinterpret_cast<const float4 *>(gt) + i);
#define CC(a, b, k_) cc[IDX((a) + ido * ((b) + cdim * (k_)))]
#define CH(a, b, k_) ch[IDX((a) + ido * ((b) + l1 * (k_)))]
    if (ido == 1) {
        // one phase: the inputs are read from the r2hc layout with their factor 2 (INMODE 2), the PM that follows the
        // O(ip^2) phase is applied on the way out (MODE 2); the result is in ch
        __syncthreads();   // sgt
        FOR_ITEMS(it, (nlb + 1) * idl1) {
            const int lb = fdiv(it, P.m_idl1), ik = it - lb * idl1;
            if (lb == nlb) {
                float s = CC(0, 0, ik);
                for (int j = 1; j < ipph; ++j) s += 2 * CC(0, 2 * j - 1, ik);
                ch[IDX(ik)] = s;
            } else generic_block_lb<2, 2>(LB, c, cc, ch, sgt, ip, ipph, idl1, lb, ik, nz);
        }
        return true;
    }
    FOR_ITEMS(it, l1 * ido) {
        const int k = fdiv(it, P.m_ido), i = it - k * ido;
        CH(i, k, 0) = CC(i, 0, k);
    }
    FOR_ITEMS(it, (ipph - 1) * l1) {
        const int jj = fdiv(it, P.m_l1), k = it - jj * l1, j = jj + 1, jc = ip - j, j2 = 2 * j - 1;
        CH(0, k, j) = 2 * CC(ido - 1, j2, k);
        CH(0, k, jc) = 2 * CC(0, j2 + 1, k);
    }
    if (ido != 1) {
        const int ni = (ido - 1) >> 1;
        FOR_ITEMS(it, (ipph - 1) * l1 * ni) {
            const int jj = fdiv(it, P.m_l1ni), rem = it - jj * (l1 * ni);
            const int k = fdiv(rem, P.m_ni), i = 1 + 2 * (rem - k * ni), ic = ido - i - 2;
            const int j = jj + 1, jc = ip - j, j2 = 2 * j - 1;
            CH(i, k, j) = CC(i, j2 + 1, k) + CC(ic, j2, k);
            CH(i, k, jc) = CC(i, j2 + 1, k) - CC(ic, j2, k);
            CH(i + 1, k, j) = CC(i + 1, j2 + 1, k) - CC(ic + 1, j2, k);
            CH(i + 1, k, jc) = CC(i + 1, j2 + 1, k) + CC(ic + 1, j2, k);
        }
    }
    __syncthreads();
    // C2(ik, l >= 1) from CH2; the l == 0 item forms CH2(ik,0) + sum_j CH2(ik,j) and parks it in C2(ik,0) (cc's slot 0
    // is free) because the other items of this phase still read the old CH2(ik,0)
    FOR_ITEMS(it, (nlb + 1) * idl1) {
        const int lb = fdiv(it, P.m_idl1), ik = it - lb * idl1;
        if (lb == nlb) {
            float s = CH2(ik, 0);
            for (int j = 1; j < ipph; ++j) s += CH2(ik, j);
            C2(ik, 0) = s;
        } else generic_block_lb<0>(LB, c, ch, cc, sgt, ip, ipph, idl1, lb, ik, nz);
    }
    __syncthreads();
    FOR_ITEMS(ik, idl1) CH2(ik, 0) = C2(ik, 0);
    FOR_ITEMS(it, (ipph - 1) * l1) {
        const int jj = fdiv(it, P.m_l1), k = it - jj * l1, j = jj + 1, jc = ip - j;
        PM(CH(0, k, jc), CH(0, k, j), C1(0, k, j), C1(0, k, jc))
    }
    if (ido != 1) {
        const int ni = (ido - 1) >> 1;
        FOR_ITEMS(it, (ipph - 1) * l1 * ni) {
            const int jj = fdiv(it, P.m_l1ni), rem = it - jj * (l1 * ni);
            const int k = fdiv(rem, P.m_ni), i = 1 + 2 * (rem - k * ni);
            const int j = jj + 1, jc = ip - j;
            const float a0 = C1(i, k, j) - C1(i + 1, k, jc);      // CH(i  ,k,j )
            const float a1 = C1(i, k, j) + C1(i + 1, k, jc);      // CH(i  ,k,jc)
            const float b0 = C1(i + 1, k, j) + C1(i, k, jc);      // CH(i+1,k,j )
            const float b1 = C1(i + 1, k, j) - C1(i, k, jc);      // CH(i+1,k,jc)
            const int idij = (j - 1) * (ido - 1) + (i - 1), idij2 = (jc - 1) * (ido - 1) + (i - 1);
            const float w0 = __ldg(wa + idij), w1 = __ldg(wa + idij + 1), v0 = __ldg(wa + idij2), v1 = __ldg(wa + idij2 + 1);
            CH(i, k, j) = w0 * a0 - w1 * b0;
            CH(i + 1, k, j) = w0 * b0 + w1 * a0;
            CH(i, k, jc) = v0 * a1 - v1 * b1;
            CH(i + 1, k, jc) = v0 * b1 + v1 * a1;
        }
    }
    return true;
#undef CC
#undef CH
#undef C1
#undef C2
#undef CH2
}
#undef WA

// ================================================================ complex passes (mirror of cfftp pass*)
// Complex arithmetic on packed f32x2 registers.  Every component sees the same IEEE operations in the same order as the
// scalar expressions of pocketfft's cmplx<float> (one rounding per product, one per sum; a - b is a + (-b) and (-a) * b is
// -(a * b) exactly), at half the instruction count.  Products are formed as fma(a, b, -0.0) with the -0.0 read from
// constant memory, which ptxas cannot fold: a plain packed multiply followed by a packed add would be re-contracted into
// FFMA2 even under -fmad=false (see mul2x above).
__constant__ float2 c_nz2 = {-0.0f, -0.0f};
__device__ __forceinline__ float2 p_mul(float2 a, float2 b) { return __ffma2_rn(a, b, c_nz2); }
__device__ __forceinline__ float2 p_muls(float s_, float2 a) { return __ffma2_rn(make_float2(s_, s_), a, c_nz2); }
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
template <bool FWD>
__device__ __forceinline__ float2 smul(float2 v1, float2 v2)   // FWD ? v1 * conj(v2) : v1 * v2
{
    // FWD: (v1.x v2.x + v1.y v2.y, v1.y v2.x - v1.x v2.y);  !FWD: (v1.x v2.x - v1.y v2.y, v1.x v2.y + v1.y v2.x)
    const float2 p = p_mul(v1, make_float2(v2.x, v2.x));
    const float2 q = p_mul(FWD ? make_float2(v1.y, -v1.x) : make_float2(-v1.y, v1.x), make_float2(v2.y, v2.y));
    return __fadd2_rn(p, q);
}
template <bool FWD>
__device__ __forceinline__ float2 rotx90(float2 a) { return FWD ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }
template <bool FWD>
__device__ __forceinline__ float2 rotx45(float2 a)
{
    const float hsqt2 = 0.707106781186547524400844362104849f;
    // FWD: h (a.x + a.y, a.y - a.x);  !FWD: h (a.x - a.y, a.y + a.x)
    return p_muls(hsqt2, __fadd2_rn(a, FWD ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x)));
}
template <bool FWD>
__device__ __forceinline__ float2 rotx135(float2 a)
{
    const float hsqt2 = 0.707106781186547524400844362104849f;
    // FWD: h (a.y - a.x, -a.x - a.y);  !FWD: h (-a.x - a.y, a.x - a.y)
    return p_muls(hsqt2, FWD ? __fadd2_rn(make_float2(a.y, -a.x), make_float2(-a.x, -a.y))
                             : __fadd2_rn(make_float2(-a.x, a.x), make_float2(-a.y, -a.y)));
}
#define CPM(a, b, cc_, d) { a = c_add(cc_, d); b = c_sub(cc_, d); }
#define CWA(x, i) __ldg(wa + (i) - 1 + (x) * (ido - 1))

// one complex pass of radix IP over `ninst` transforms of length n2 per row; element e of instance q at (q*n2 + e)
template <int IP, bool FWD>
__device__ void cpass(const Ctx &c, int ido, int l1, const float2 *cc0, float2 *ch0, const float2 *wa, int ninst, int n2,
                      unsigned m_ido, unsigned m_per)
{
    const int per = l1 * ido;
    FOR_ITEMS(it, ninst * per) {
        // (instance, k, i) of the butterfly: divisions by the pass constants through their magic multipliers
        const int q = ninst == 1 ? 0 : fdiv(it, m_per), rem = it - q * per;
        const int k = fdiv(rem, m_ido), i = rem - k * ido;
        const float2 *cc = cc0 + (size_t)q * n2 * c.GP;
        float2 *ch = ch0 + (size_t)q * n2 * c.GP;
#define CCC(a, b, k_) cc[IDX((a) + ido * ((b) + IP * (k_)))]
#define CCH(a, k_, b) ch[IDX((a) + ido * ((k_) + l1 * (b)))]
        if (IP == 2) {
            const float2 a = CCC(i, 0, k), b = CCC(i, 1, k);
            CCH(i, k, 0) = c_add(a, b);
            CCH(i, k, 1) = i == 0 ? c_sub(a, b) : smul<FWD>(c_sub(a, b), CWA(0, i));
        } else if (IP == 3) {
            const float tw1r = -0.5f, tw1i = (FWD ? -1.f : 1.f) * 0.8660254037844386467637231707529362f;
            const float2 t0 = CCC(i, 0, k);
            float2 t1, t2;
            CPM(t1, t2, CCC(i, 1, k), CCC(i, 2, k))
            CCH(i, k, 0) = c_add(t0, t1);
            const float2 ca = c_add(t0, p_muls(tw1r, t1));
            const float2 cb = p_muls(tw1i, make_float2(-t2.y, t2.x));
            if (i == 0) { CPM(CCH(0, k, 1), CCH(0, k, 2), ca, cb) }
            else {
                CCH(i, k, 1) = smul<FWD>(c_add(ca, cb), CWA(0, i));
                CCH(i, k, 2) = smul<FWD>(c_sub(ca, cb), CWA(1, i));
            }
        } else if (IP == 4) {
            float2 t1, t2, t3, t4;
            CPM(t2, t1, CCC(i, 0, k), CCC(i, 2, k))
            CPM(t3, t4, CCC(i, 1, k), CCC(i, 3, k))
            t4 = rotx90<FWD>(t4);
            if (i == 0) {
                CPM(CCH(0, k, 0), CCH(0, k, 2), t2, t3)
                CPM(CCH(0, k, 1), CCH(0, k, 3), t1, t4)
            } else {
                CCH(i, k, 0) = c_add(t2, t3);
                CCH(i, k, 1) = smul<FWD>(c_add(t1, t4), CWA(0, i));
                CCH(i, k, 2) = smul<FWD>(c_sub(t2, t3), CWA(1, i));
                CCH(i, k, 3) = smul<FWD>(c_sub(t1, t4), CWA(2, i));
            }
        } else if (IP == 5) {
            const float tw1r = 0.3090169943749474241022934171828191f,
                        tw1i = (FWD ? -1.f : 1.f) * 0.9510565162951535721164393333793821f,
                        tw2r = -0.8090169943749474241022934171828191f,
                        tw2i = (FWD ? -1.f : 1.f) * 0.5877852522924731291687059546390728f;
            const float2 t0 = CCC(i, 0, k);
            float2 t1, t2, t3, t4;
            CPM(t1, t4, CCC(i, 1, k), CCC(i, 4, k))
            CPM(t2, t3, CCC(i, 2, k), CCC(i, 3, k))
            CCH(i, k, 0) = c_add(c_add(t0, t1), t2);
#define STEP5(u1, u2, twar, twbr, twai, twbi)                                           \
            {                                                                           \
                const float2 ca = c_add(c_add(t0, p_muls(twar, t1)), p_muls(twbr, t2)); \
                const float2 u_ = c_add(p_muls((twai), t4), p_muls((twbi), t3));        \
                const float2 cb = make_float2(-u_.y, u_.x);                             \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }               \
                else {                                                                  \
                    CCH(i, k, u1) = smul<FWD>(c_add(ca, cb), CWA(u1 - 1, i));           \
                    CCH(i, k, u2) = smul<FWD>(c_sub(ca, cb), CWA(u2 - 1, i));           \
                }                                                                       \
            }
            STEP5(1, 4, tw1r, tw2r, +tw1i, +tw2i)
            STEP5(2, 3, tw2r, tw1r, +tw2i, -tw1i)
#undef STEP5
        } else if (IP == 7) {
            const float tw1r = 0.6234898018587335305250048840042398f,
                        tw1i = (FWD ? -1.f : 1.f) * 0.7818314824680298087084445266740578f,
                        tw2r = -0.2225209339563144042889025644967948f,
                        tw2i = (FWD ? -1.f : 1.f) * 0.9749279121818236070181316829939312f,
                        tw3r = -0.9009688679024191262361023195074451f,
                        tw3i = (FWD ? -1.f : 1.f) * 0.433883739117558120475768332848359f;
            const float2 t1 = CCC(i, 0, k);
            float2 t2, t3, t4, t5, t6, t7;
            CPM(t2, t7, CCC(i, 1, k), CCC(i, 6, k))
            CPM(t3, t6, CCC(i, 2, k), CCC(i, 5, k))
            CPM(t4, t5, CCC(i, 3, k), CCC(i, 4, k))
            CCH(i, k, 0) = c_add(c_add(c_add(t1, t2), t3), t4);
#define STEP7(u1, u2, x1, x2, x3, y1, y2, y3)                                           \
            {                                                                           \
                const float2 ca = c_add(c_add(c_add(t1, p_muls(x1, t2)), p_muls(x2, t3)), p_muls(x3, t4)); \
                const float2 u_ = c_add(c_add(p_muls((y1), t7), p_muls((y2), t6)), p_muls((y3), t5));     \
                const float2 cb = make_float2(-u_.y, u_.x);                             \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }               \
                else {                                                                  \
                    CCH(i, k, u1) = smul<FWD>(c_add(ca, cb), CWA(u1 - 1, i));           \
                    CCH(i, k, u2) = smul<FWD>(c_sub(ca, cb), CWA(u2 - 1, i));           \
                }                                                                       \
            }
            STEP7(1, 6, tw1r, tw2r, tw3r, +tw1i, +tw2i, +tw3i)
            STEP7(2, 5, tw2r, tw3r, tw1r, +tw2i, -tw3i, -tw1i)
            STEP7(3, 4, tw3r, tw1r, tw2r, +tw3i, -tw1i, +tw2i)
#undef STEP7
        } else if (IP == 8) {
            float2 a0, a1, a2, a3, a4, a5, a6, a7;
            CPM(a1, a5, CCC(i, 1, k), CCC(i, 5, k))
            CPM(a3, a7, CCC(i, 3, k), CCC(i, 7, k))
            { const float2 s = c_add(a1, a3), d = c_sub(a1, a3); a1 = s; a3 = d; }
            a3 = rotx90<FWD>(a3);
            a7 = rotx90<FWD>(a7);
            { const float2 s = c_add(a5, a7), d = c_sub(a5, a7); a5 = s; a7 = d; }
            a5 = rotx45<FWD>(a5);
            a7 = rotx135<FWD>(a7);
            CPM(a0, a4, CCC(i, 0, k), CCC(i, 4, k))
            CPM(a2, a6, CCC(i, 2, k), CCC(i, 6, k))
            if (i == 0) {
                CPM(CCH(0, k, 0), CCH(0, k, 4), c_add(a0, a2), a1)
                CPM(CCH(0, k, 2), CCH(0, k, 6), c_sub(a0, a2), a3)
                a6 = rotx90<FWD>(a6);
                CPM(CCH(0, k, 1), CCH(0, k, 5), c_add(a4, a6), a5)
                CPM(CCH(0, k, 3), CCH(0, k, 7), c_sub(a4, a6), a7)
            } else {
                { const float2 s = c_add(a0, a2), d = c_sub(a0, a2); a0 = s; a2 = d; }
                CCH(i, k, 0) = c_add(a0, a1);
                CCH(i, k, 4) = smul<FWD>(c_sub(a0, a1), CWA(3, i));
                CCH(i, k, 2) = smul<FWD>(c_add(a2, a3), CWA(1, i));
                CCH(i, k, 6) = smul<FWD>(c_sub(a2, a3), CWA(5, i));
                a6 = rotx90<FWD>(a6);
                { const float2 s = c_add(a4, a6), d = c_sub(a4, a6); a4 = s; a6 = d; }
                CCH(i, k, 1) = smul<FWD>(c_add(a4, a5), CWA(0, i));
                CCH(i, k, 5) = smul<FWD>(c_sub(a4, a5), CWA(4, i));
                CCH(i, k, 3) = smul<FWD>(c_add(a6, a7), CWA(2, i));
                CCH(i, k, 7) = smul<FWD>(c_sub(a6, a7), CWA(6, i));
            }
        } else if (IP == 11) {
            const float tw1r = 0.8412535328311811688618116489193677f,
                        tw1i = (FWD ? -1.f : 1.f) * 0.5406408174555975821076359543186917f,
                        tw2r = 0.4154150130018864255292741492296232f,
                        tw2i = (FWD ? -1.f : 1.f) * 0.9096319953545183714117153830790285f,
                        tw3r = -0.1423148382732851404437926686163697f,
                        tw3i = (FWD ? -1.f : 1.f) * 0.9898214418809327323760920377767188f,
                        tw4r = -0.6548607339452850640569250724662936f,
                        tw4i = (FWD ? -1.f : 1.f) * 0.7557495743542582837740358439723444f,
                        tw5r = -0.9594929736144973898903680570663277f,
                        tw5i = (FWD ? -1.f : 1.f) * 0.2817325568414296977114179153466169f;
            const float2 t1 = CCC(i, 0, k);
            float2 t2, t3, t4, t5, t6, t7, t8, t9, t10, t11;
            CPM(t2, t11, CCC(i, 1, k), CCC(i, 10, k))
            CPM(t3, t10, CCC(i, 2, k), CCC(i, 9, k))
            CPM(t4, t9, CCC(i, 3, k), CCC(i, 8, k))
            CPM(t5, t8, CCC(i, 4, k), CCC(i, 7, k))
            CPM(t6, t7, CCC(i, 5, k), CCC(i, 6, k))
            CCH(i, k, 0) = c_add(c_add(c_add(c_add(c_add(t1, t2), t3), t4), t5), t6);
#define STEP11(u1, u2, x1, x2, x3, x4, x5, y1, y2, y3, y4, y5)                                      \
            {                                                                                       \
                const float2 ca = c_add(c_add(c_add(c_add(c_add(t1, p_muls(x1, t2)), p_muls(x2, t3)), p_muls(x3, t4)), \
                                              p_muls(x4, t5)), p_muls(x5, t6));                     \
                const float2 u_ = c_add(c_add(c_add(c_add(p_muls((y1), t11), p_muls((y2), t10)), p_muls((y3), t9)), \
                                              p_muls((y4), t8)), p_muls((y5), t7));                 \
                const float2 cb = make_float2(-u_.y, u_.x);                                         \
                if (i == 0) { CPM(CCH(0, k, u1), CCH(0, k, u2), ca, cb) }                           \
                else {                                                                              \
                    CCH(i, k, u1) = smul<FWD>(c_add(ca, cb), CWA(u1 - 1, i));                       \
                    CCH(i, k, u2) = smul<FWD>(c_sub(ca, cb), CWA(u2 - 1, i));                       \
                }                                                                                   \
            }
            STEP11(1, 10, tw1r, tw2r, tw3r, tw4r, tw5r, +tw1i, +tw2i, +tw3i, +tw4i, +tw5i)
            STEP11(2, 9, tw2r, tw4r, tw5r, tw3r, tw1r, +tw2i, +tw4i, -tw5i, -tw3i, -tw1i)
            STEP11(3, 8, tw3r, tw5r, tw2r, tw1r, tw4r, +tw3i, -tw5i, -tw2i, +tw1i, +tw4i)
            STEP11(4, 7, tw4r, tw3r, tw1r, tw5r, tw2r, +tw4i, -tw3i, +tw1i, +tw5i, -tw2i)
            STEP11(5, 6, tw5r, tw1r, tw4r, tw2r, tw3r, +tw5i, -tw1i, +tw4i, -tw2i, +tw3i)
#undef STEP11
        }
#undef CCC
#undef CCH
    }
}

struct XBlue {
    int ip, n2, nf;
    int fct[12], tw[12];   // complex sub-plan: factors in pass order, twiddle offsets (floats) into the table
    unsigned m_ido[12], m_per[12], m_n2;   // fdiv magics of ido, l1 * ido per pass and of n2
    int bk, bkf;           // offsets (floats) of bk[ip] and bkf[n2/2+1] (complex)
    int inst;              // Bluestein transforms evaluated concurrently per row
};
struct XArgs {
    B2sImg img;
    const float *g;      // notch over packed positions
    const float *tab;    // twiddles etc.
    int n, nseq, along_cols;
    int G, lgG;
    int nf;
    XPass fwd[12], bwd[12];
    XBlue blue;
    float fct;           // 1/n
    float negzero;       // -0.0f, deliberately a run-time value (see mul2x)
    int gt_max;          // float2 entries reserved for the generic-radix table
    int groups_per_plane;
};

// generic complex radix (cfftp::passg / ducc0 cfftpg), one transform per row; the result ends in cc.
// wal: shared-memory copy of csarr with the direction's sign (ip entries)
template <bool FWD>
__device__ void cpassg(const Ctx &c, int ido, int ip, int l1, float2 *cc, float2 *ch, const float2 *wa, const float2 *csarr,
                       float2 *wal)
{
    const int cdim = ip, ipph = (ip + 1) / 2, idl1 = ido * l1;
#define GCH(a, b, k_) ch[IDX((a) + ido * ((b) + l1 * (k_)))]
#define GCC(a, b, k_) cc[IDX((a) + ido * ((b) + cdim * (k_)))]
#define GCX(a, b, k_) cc[IDX((a) + ido * ((b) + l1 * (k_)))]
#define GCX2(a, b) cc[IDX((a) + idl1 * (b))]
#define GCH2(a, b) ch[IDX((a) + idl1 * (b))]
    for (int i = threadIdx.x; i < ip; i += kXT) {
        const float2 w = __ldg(csarr + i);
        wal[i] = i == 0 ? make_float2(1.f, 0.f) : make_float2(w.x, FWD ? -w.y : w.y);
    }
    FOR_ITEMS(it, ipph * idl1) {
        const int j = it / idl1, rem = it - j * idl1;
        const int k = rem / ido, i = rem - k * ido;
        if (j == 0) GCH(i, k, 0) = GCC(i, 0, k);
        else { const int jc = ip - j; CPM(GCH(i, k, j), GCH(i, k, jc), GCC(i, j, k), GCC(i, jc, k)) }
    }
    __syncthreads();
    FOR_ITEMS(it, ipph * idl1) {
        const int l = it / idl1, ik = it - l * idl1;
        if (l == 0) {
            float2 tmp = GCH2(ik, 0);
            for (int j = 1; j < ipph; ++j) { const float2 v = GCH2(ik, j); tmp.x += v.x; tmp.y += v.y; }
            GCX2(ik, 0) = tmp;
        } else {
            const int lc = ip - l;
            const float2 h0 = GCH2(ik, 0), h1 = GCH2(ik, 1), h2 = GCH2(ik, 2), g1 = GCH2(ik, ip - 1), g2 = GCH2(ik, ip - 2);
            const float2 w1 = wal[l], w2 = wal[2 * l];
            float2 a, b;
            a.x = h0.x + w1.x * h1.x + w2.x * h2.x;
            a.y = h0.y + w1.x * h1.y + w2.x * h2.y;
            b.x = -(w1.y * g1.y + w2.y * g2.y);
            b.y = w1.y * g1.x + w2.y * g2.x;
            int iwal = 2 * l;
            int j = 3, jc = ip - 3;
            for (; j + 1 < ipph; j += 2, jc -= 2) {
                iwal += l; if (iwal > ip) iwal -= ip;
                const float2 xw = wal[iwal];
                iwal += l; if (iwal > ip) iwal -= ip;
                const float2 xw2 = wal[iwal];
                const float2 p0 = GCH2(ik, j), p1 = GCH2(ik, j + 1), q0 = GCH2(ik, jc), q1 = GCH2(ik, jc - 1);
                a.x += p0.x * xw.x + p1.x * xw2.x;
                a.y += p0.y * xw.x + p1.y * xw2.x;
                b.x -= q0.y * xw.y + q1.y * xw2.y;
                b.y += q0.x * xw.y + q1.x * xw2.y;
            }
            for (; j < ipph; ++j, --jc) {
                iwal += l; if (iwal > ip) iwal -= ip;
                const float2 xw = wal[iwal];
                const float2 p0 = GCH2(ik, j), q0 = GCH2(ik, jc);
                a.x += p0.x * xw.x;
                a.y += p0.y * xw.x;
                b.x -= q0.y * xw.y;
                b.y += q0.x * xw.y;
            }
            GCX2(ik, l) = a;
            GCX2(ik, lc) = b;
        }
    }
    __syncthreads();
    // shuffling and twiddling, in place
    FOR_ITEMS(it, (ipph - 1) * idl1) {
        const int jj = it / idl1, rem = it - jj * idl1;
        const int k = rem / ido, i = rem - k * ido;
        const int j = jj + 1, jc = ip - j;
        const float2 t1 = GCX(i, k, j), t2 = GCX(i, k, jc);
        float2 x1, x2;
        CPM(x1, x2, t1, t2)
        if (i == 0) { GCX(i, k, j) = x1; GCX(i, k, jc) = x2; }
        else {
            GCX(i, k, j) = smul<FWD>(x1, __ldg(wa + (j - 1) * (ido - 1) + i - 1));
            GCX(i, k, jc) = smul<FWD>(x2, __ldg(wa + (jc - 1) * (ido - 1) + i - 1));
        }
    }
#undef GCH
#undef GCC
#undef GCX
#undef GCX2
#undef GCH2
}

template <bool FWD>
__device__ void cblue(const Ctx &c, const XBlue &b, const float *tab, int ido, int l1, float2 *cc, float2 *ch,
                      const float2 *wa, float2 *X0, float2 *X1);

// cs: per-factor offsets (floats) of csarr for generic radices (may be null when every factor is hard-coded);
// blue / BX0 / BX1: Bluestein sub-plan and work buffers for a prime factor >= 110 (complexify plans only)
// OUTER: the plan may contain generic / Bluestein factors (complexify plans); the inner 11-smooth plans of a Bluestein
// convolution are instantiated with OUTER = false, which also keeps the call graph free of recursion (everything inlines)
template <bool FWD, bool OUTER = false>
__device__ __forceinline__ void cfft_all(const Ctx &c, const XBlue &b, const float *tab, float2 *&cur, float2 *&nxt, int ninst,
                                         const int *cs = nullptr, float2 *wal = nullptr, const XBlue *blue = nullptr,
                                         float2 *BX0 = nullptr, float2 *BX1 = nullptr)
{
    int l1 = 1;
    for (int f = 0; f < b.nf; ++f) {
        const int ip = b.fct[f], ido = b.n2 / (l1 * ip);
        const float2 *wa = reinterpret_cast<const float2 *>(tab + b.tw[f]);
        bool swap = true;
        switch (ip) {
        case 2: cpass<2, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 3: cpass<3, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 4: cpass<4, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 5: cpass<5, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 7: cpass<7, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 8: cpass<8, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        case 11: cpass<11, FWD>(c, ido, l1, cur, nxt, wa, ninst, b.n2, b.m_ido[f], b.m_per[f]); break;
        default:
            if (OUTER) {
                if (ip >= 110) { cblue<FWD>(c, *blue, tab, ido, l1, cur, nxt, wa, BX0, BX1); swap = l1 > 1; }
                else { cpassg<FWD>(c, ido, ip, l1, cur, nxt, wa, reinterpret_cast<const float2 *>(tab + cs[f]), wal); swap = false; }
            }
            break;
        }
        __syncthreads();
        if (swap) { float2 *t = cur; cur = nxt; nxt = t; }
        l1 *= ip;
    }
}

// Bluestein pass of a complex plan (ducc0 cfftpblue, prime radix >= 110): every (k, i) column is one transform of length
// ip evaluated as a cyclic convolution of length n2; the inter-pass twiddle is merged into the final multiplication by
// b_k.  l1 > 1: result in ch; l1 == 1: written back into cc.
template <bool FWD>
__device__ void cblue(const Ctx &c, const XBlue &b, const float *tab, int ido, int l1, float2 *cc, float2 *ch,
                      const float2 *wa, float2 *X0, float2 *X1)
{
    const int ip = b.ip, n2 = b.n2;
    const float2 *bk = reinterpret_cast<const float2 *>(tab + b.bk);
    const float2 *bkf = reinterpret_cast<const float2 *>(tab + b.bkf);
    const int total = l1 * ido;
#define BCC(a, m_, k_) cc[IDX((a) + ido * ((m_) + ip * (k_)))]
#define BCH(a, k_, m_) ch[IDX((a) + ido * ((k_) + l1 * (m_)))]
    for (int base = 0; base < total; base += b.inst) {
        const int ninst = min(b.inst, total - base);
        FOR_ITEMS(it, ninst * n2) {
            const int q = ninst == 1 ? 0 : fdiv(it, b.m_n2), m = it - q * n2;
            const int inst = base + q, k = inst / ido, i = inst - k * ido;
            const float2 a0 = smul<FWD>(BCC(i, 0, k), __ldg(bk));
            X0[IDX(q * n2 + m)] = m < ip ? smul<FWD>(BCC(i, m, k), __ldg(bk + m)) : make_float2(a0.x * 0.f, a0.y * 0.f);
        }
        __syncthreads();
        float2 *cur = X0, *nxt = X1;
        cfft_all<true>(c, b, tab, cur, nxt, ninst);
        FOR_ITEMS(it, ninst * n2) {
            const int q = ninst == 1 ? 0 : fdiv(it, b.m_n2), m = it - q * n2;
            const int mb = 2 * m <= n2 ? m : n2 - m;
            cur[IDX(q * n2 + m)] = smul<!FWD>(cur[IDX(q * n2 + m)], __ldg(bkf + mb));
        }
        __syncthreads();
        cfft_all<false>(c, b, tab, cur, nxt, ninst);
        FOR_ITEMS(it, ninst * ip) {
            const int q = ninst == 1 ? 0 : it / ip, m = it - q * ip;
            const int inst = base + q, k = inst / ido, i = inst - k * ido;
            float2 w = __ldg(bk + m);
            if (i != 0 && m != 0) {
                const float2 t = __ldg(wa + (i - 1) + (m - 1) * (ido - 1));
                w = make_float2(w.x * t.x - w.y * t.y, w.x * t.y + w.y * t.x);
            }
            const float2 v = smul<FWD>(cur[IDX(q * n2 + m)], w);
            if (l1 > 1) BCH(i, k, m) = v; else BCC(i, m, 0) = v;
        }
        __syncthreads();
    }
#undef BCC
#undef BCH
}

// Bluestein pass of the real transform (ducc0 rfftpblue): every (k) and (k, i) column set is one complex transform of
// length ip evaluated as a cyclic convolution of length n2.  FWD: cc -> ch in radf layout; !FWD: radb layout.
template <bool FWD>
__device__ void rblue(const Ctx &c, const XBlue &b, const float *tab, int ido, int l1, const float *cc, float *ch,
                      const float *wa, float2 *X0, float2 *X1)
{
    const int ip = b.ip, n2 = b.n2, ipph = (ip + 1) / 2;
    const float2 *bk = reinterpret_cast<const float2 *>(tab + b.bk);
    const float2 *bkf = reinterpret_cast<const float2 *>(tab + b.bkf);
    const int ni = (ido - 1) >> 1;            // i = 2, 4, ... instances per k
    const int per_k = 1 + ni;
    const int total = l1 * per_k;
#define WA(x, i) __ldg(wa + (i) + (x) * (ido - 1))
    for (int base = 0; base < total; base += b.inst) {
        const int ninst = min(b.inst, total - base);
        // ---- a_m = x_m * conj/plain(b_m), zero padded
        FOR_ITEMS(it, ninst * n2) {
            const int q = ninst == 1 ? 0 : fdiv(it, b.m_n2), m = it - q * n2;
            const int inst = base + q, k = per_k == 1 ? inst : inst / per_k, ii = inst - k * per_k, i = 2 * ii, ic = ido - i;
            float2 x0;   // element 0 of the transform's input
            if (FWD) {
#define CC(a, k_, m_) cc[IDX((a) + ido * ((k_) + l1 * (m_)))]
                x0 = ii == 0 ? make_float2(CC(0, k, 0), 0.f) : make_float2(CC(i - 1, k, 0), CC(i, k, 0));
                float2 v;
                if (m == 0) v = x0;
                else if (m < ip) {
                    if (ii == 0) v = make_float2(CC(0, k, m), 0.f);
                    else MULPM(v.x, v.y, WA(m - 1, i - 2), WA(m - 1, i - 1), CC(i - 1, k, m), CC(i, k, m))
                }
                const float2 a0 = smul<FWD>(x0, __ldg(bk));
                X0[IDX(q * n2 + m)] = m < ip ? smul<FWD>(v, __ldg(bk + m)) : make_float2(a0.x * 0.f, a0.y * 0.f);
#undef CC
            } else {
#define CC(a, m_, k_) cc[IDX((a) + ido * ((m_) + ip * (k_)))]
                x0 = ii == 0 ? make_float2(CC(0, 0, k), 0.f) : make_float2(CC(i - 1, 0, k), CC(i, 0, k));
                float2 v;
                if (m == 0) v = x0;
                else if (m < ip) {
                    const int mm = m < ipph ? m : ip - m;
                    if (ii == 0) v = make_float2(CC(ido - 1, 2 * mm - 1, k), m < ipph ? CC(0, 2 * mm, k) : -CC(0, 2 * mm, k));
                    else if (m < ipph) v = make_float2(CC(i - 1, 2 * mm, k), CC(i, 2 * mm, k));
                    else v = make_float2(CC(ic - 1, 2 * mm - 1, k), -CC(ic, 2 * mm - 1, k));
                }
                const float2 a0 = smul<FWD>(x0, __ldg(bk));
                X0[IDX(q * n2 + m)] = m < ip ? smul<FWD>(v, __ldg(bk + m)) : make_float2(a0.x * 0.f, a0.y * 0.f);
#undef CC
            }
        }
        __syncthreads();
        float2 *cur = X0, *nxt = X1;
        cfft_all<true>(c, b, tab, cur, nxt, ninst);
        // ---- convolution: multiply by the (half-stored, symmetric) transform of b
        FOR_ITEMS(it, ninst * n2) {
            const int q = ninst == 1 ? 0 : fdiv(it, b.m_n2), m = it - q * n2;
            const int mb = 2 * m <= n2 ? m : n2 - m;
            cur[IDX(q * n2 + m)] = smul<!FWD>(cur[IDX(q * n2 + m)], __ldg(bkf + mb));
        }
        __syncthreads();
        cfft_all<false>(c, b, tab, cur, nxt, ninst);
        // ---- multiply by b_k and scatter into the pass output
        FOR_ITEMS(it, ninst * ip) {
            const int q = ninst == 1 ? 0 : it / ip, m = it - q * ip;
            const int inst = base + q, k = per_k == 1 ? inst : inst / per_k, ii = inst - k * per_k, i = 2 * ii, ic = ido - i;
            const float2 res = smul<FWD>(cur[IDX(q * n2 + m)], __ldg(bk + m));
            if (FWD) {
#define CH(a, m_, k_) ch[IDX((a) + ido * ((m_) + ip * (k_)))]
                if (ii == 0) {
                    if (m == 0) CH(0, 0, k) = res.x;
                    else if (m < ipph) { CH(ido - 1, 2 * m - 1, k) = res.x; CH(0, 2 * m, k) = res.y; }
                } else {
                    if (m == 0) { CH(i - 1, 0, k) = res.x; CH(i, 0, k) = res.y; }
                    else if (m < ipph) { CH(i - 1, 2 * m, k) = res.x; CH(i, 2 * m, k) = res.y; }
                    else { const int mm = ip - m; CH(ic - 1, 2 * mm - 1, k) = res.x; CH(ic, 2 * mm - 1, k) = -res.y; }
                }
#undef CH
            } else {
#define CH(a, k_, m_) ch[IDX((a) + ido * ((k_) + l1 * (m_)))]
                if (ii == 0) CH(0, k, m) = res.x;
                else if (m == 0) { CH(i - 1, k, 0) = res.x; CH(i, k, 0) = res.y; }
                else MULPM(CH(i, k, m), CH(i - 1, k, m), WA(m - 1, i - 2), WA(m - 1, i - 1), res.y, res.x)
#undef CH
            }
        }
        __syncthreads();
    }
#undef WA
}

__global__ void __launch_bounds__(kXT, 2) k_notch_exact(const __grid_constant__ XArgs a)
{
    extern __shared__ __align__(16) float xs[];
    const int n = a.n, G = a.G, GP = G;   // unpadded: 32/G consecutive elements x G rows hit 32 different banks
    float *A = xs, *B = A + n * GP;
    float2 *sgt = reinterpret_cast<float2 *>(xs + ((2 * n * GP + 3) & ~3));   // generic-radix (cos, sin) table of the running pass
    float2 *X0 = sgt + a.gt_max;                            // Bluestein work buffers
    const float2 nz = make_float2(a.negzero, a.negzero);
    float2 *X1 = X0 + (size_t)a.blue.inst * a.blue.n2 * GP;

    Ctx c;
    c.r = threadIdx.x & (G - 1);
    c.item0 = threadIdx.x >> a.lgG;
    c.istride = kXT >> a.lgG;
    c.GP = GP;

    float *plane = a.img.ptr + (size_t)blockIdx.y * a.img.plane_stride;
    for (int grp = blockIdx.x; grp < a.groups_per_plane; grp += gridDim.x) {
        const int s0 = grp * G;
        const int ns = min(G, a.nseq - s0);
        __syncthreads();
        // ---- gather (rows beyond the sub-band are zero sequences)
        // a warp covers 32/G consecutive elements of G rows: each row contributes one 16..128-byte run
        if (!a.along_cols) {
            for (int idx = threadIdx.x; idx < G * n; idx += kXT) {
                const int e = idx >> a.lgG, rr = idx & (G - 1);
                A[idx] = rr < ns ? plane[(size_t)(s0 + rr) * a.img.pitch + e] : 0.f;
            }
        } else {
            for (int idx = threadIdx.x; idx < G * n; idx += kXT) {
                const int e = idx >> a.lgG, rr = idx & (G - 1);
                A[idx] = rr < ns ? plane[(size_t)e * a.img.pitch + s0 + rr] : 0.f;
            }
        }
        __syncthreads();
        float *p1 = A, *p2 = B;
        // ---- forward: rfftp::exec(r2hc = true)
        for (int f = 0; f < a.nf; ++f) {
            const XPass &p = a.fwd[f];
            const float *wa = a.tab + p.tw;
            bool swap = true;
            switch (p.kind) {
            case 2: radf2(c, p.ido, p.l1, p1, p2, wa); break;
            case 3: radf3(c, p.ido, p.l1, p1, p2, wa); break;
            case 4: radf4(c, p.ido, p.l1, p1, p2, wa); break;
            case 5: radf5(c, p.ido, p.l1, p1, p2, wa); break;
            case 6: radfg(c, p, p1, p2, wa, reinterpret_cast<const float2 *>(a.tab + p.cs), sgt, nz); break;   // result in p2
            default: rblue<true>(c, a.blue, a.tab, p.ido, p.l1, p1, p2, wa, X0, X1); break;
            }
            __syncthreads();
            if (swap) { float *t = p1; p1 = p2; p2 = t; }
        }
        // ---- notch on packed positions (core.py:752: spec *= g)
        for (int idx = threadIdx.x; idx < G * n; idx += kXT) {
            const int e = idx >> a.lgG, rr = idx & (G - 1);
            p1[idx] = p1[idx] * __ldg(a.g + e);
        }
        __syncthreads();
        // ---- backward: rfftp::exec(r2hc = false)
        for (int f = 0; f < a.nf; ++f) {
            const XPass &p = a.bwd[f];
            const float *wa = a.tab + p.tw;
            bool swap = true;
            switch (p.kind) {
            case 2: radb2(c, p.ido, p.l1, p1, p2, wa); break;
            case 3: radb3(c, p.ido, p.l1, p1, p2, wa); break;
            case 4: radb4(c, p.ido, p.l1, p1, p2, wa); break;
            case 5: radb5(c, p.ido, p.l1, p1, p2, wa); break;
            case 6: swap = radbg(c, p, p1, p2, wa, reinterpret_cast<const float2 *>(a.tab + p.cs), sgt, nz); break;
            default: rblue<false>(c, a.blue, a.tab, p.ido, p.l1, p1, p2, wa, X0, X1); break;
            }
            __syncthreads();
            if (swap) { float *t = p1; p1 = p2; p2 = t; }
        }
        // ---- scale by 1/n (copy_and_norm) and scatter
        if (!a.along_cols) {
            for (int idx = threadIdx.x; idx < G * n; idx += kXT) {
                const int e = idx >> a.lgG, rr = idx & (G - 1);
                if (rr < ns) plane[(size_t)(s0 + rr) * a.img.pitch + e] = a.fct * p1[idx];
            }
        } else {
            for (int idx = threadIdx.x; idx < G * n; idx += kXT) {
                const int e = idx >> a.lgG, rr = idx & (G - 1);
                if (rr < ns) plane[(size_t)e * a.img.pitch + s0 + rr] = a.fct * p1[idx];
            }
        }
    }
}

// ---- even lengths > 1000: ducc0 rfftp_complexify -------------------------------------------------------------------
// The real transform of length N runs as one complex transform of length N/2 on (x[2m], x[2m+1]) plus a butterfly with
// the N-point roots; the inverse undoes it.  The post-processing of the forward transform, the notch and the
// pre-processing of the inverse touch the same four packed positions per (i, N/2-i) pair, so they are one fused step
// here and the packed spectrum never exists in memory.
struct XCArgs {
    B2sImg img;
    const float *g;
    const float *tab;
    int n, nseq, along_cols;
    int G, lgG;
    XBlue cp;            // complex plan of length n/2 (ip / bk / bkf / inst unused)
    XBlue blue;          // Bluestein sub-plan of the (single) prime factor >= 110 of n/2, if any
    int cs[12];          // csarr offsets of generic factors
    int roots;           // offset (floats) of exp(2 pi i k / n), k <= n/4
    int wal_max;         // largest generic radix
    float fct;
    int groups_per_plane;
};

__global__ void __launch_bounds__(kXT, 2) k_notch_cplx(const __grid_constant__ XCArgs a)
{
    extern __shared__ __align__(16) float xs[];
    const int n = a.n, h = n >> 1, G = a.G;
    float2 *X0 = reinterpret_cast<float2 *>(xs), *X1 = X0 + (size_t)h * G;
    float2 *wal = X1 + (size_t)h * G;
    float2 *BX0 = wal + a.wal_max + 2, *BX1 = BX0 + (size_t)a.blue.inst * a.blue.n2 * G;
    Ctx c;
    c.r = threadIdx.x & (G - 1);
    c.item0 = threadIdx.x >> a.lgG;
    c.istride = kXT >> a.lgG;
    c.GP = G;
    const float2 *roots = reinterpret_cast<const float2 *>(a.tab + a.roots);
    float *plane = a.img.ptr + (size_t)blockIdx.y * a.img.plane_stride;
    for (int grp = blockIdx.x; grp < a.groups_per_plane; grp += gridDim.x) {
        const int s0 = grp * G;
        const int ns = min(G, a.nseq - s0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < G * h; idx += kXT) {
            const int m = idx >> a.lgG, rr = idx & (G - 1);
            float2 v = make_float2(0.f, 0.f);
            if (rr < ns) {
                if (!a.along_cols) v = *reinterpret_cast<const float2 *>(plane + (size_t)(s0 + rr) * a.img.pitch + 2 * m);
                else v = make_float2(plane[(size_t)(2 * m) * a.img.pitch + s0 + rr], plane[(size_t)(2 * m + 1) * a.img.pitch + s0 + rr]);
            }
            X0[idx] = v;
        }
        __syncthreads();
        float2 *cur = X0, *nxt = X1;
        cfft_all<true, true>(c, a.cp, a.tab, cur, nxt, 1, a.cs, wal, &a.blue, BX0, BX1);
        // ---- forward post-processing -> notch on the four packed positions -> inverse pre-processing
        FOR_ITEMS(i, h / 2 + 1) {
            const int xi = h - i;
            if (i == 0) {
                const float2 r0 = cur[IDX(0)];
                const float c0 = (r0.x + r0.y) * __ldg(a.g), cn = (r0.x - r0.y) * __ldg(a.g + n - 1);
                cur[IDX(0)] = make_float2(c0 + cn, c0 - cn);
                continue;
            }
            const float2 ri = cur[IDX(i)], rx = cur[IDX(xi)];
            const float2 w = __ldg(roots + i);
            const float2 xe = make_float2(ri.x + rx.x, ri.y - rx.y);
            const float2 t = make_float2(ri.y + rx.y, rx.x - ri.x);
            const float2 xo = make_float2(t.x * w.x + t.y * w.y, t.y * w.x - t.x * w.y);
            float ca = 0.5f * (xe.x + xo.x), cb = 0.5f * (xe.y + xo.y);        // packed 2i-1, 2i
            const float cc_ = 0.5f * (xe.x - xo.x), cd = 0.5f * (xo.y - xe.y);  // packed 2xi-1, 2xi
            if (i == xi) { ca = cc_; cb = cd; }                                // the later assignment wins in the reference loop
            const float g1 = __ldg(a.g + 2 * i - 1), g2 = __ldg(a.g + 2 * i), g3 = __ldg(a.g + 2 * xi - 1), g4 = __ldg(a.g + 2 * xi);
            const float2 t1 = make_float2(ca * g1, cb * g2);
            const float2 t2 = make_float2(cc_ * g3, -(cd * g4));
            const float2 ye = make_float2(t1.x + t2.x, t1.y + t2.y);
            const float2 d = make_float2(t1.x - t2.x, t1.y - t2.y);
            const float2 yo = make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
            if (i != xi) cur[IDX(i)] = make_float2(ye.x - yo.y, ye.y + yo.x);
            cur[IDX(xi)] = make_float2(ye.x + yo.y, -ye.y + yo.x);
        }
        __syncthreads();
        cfft_all<false, true>(c, a.cp, a.tab, cur, nxt, 1, a.cs, wal, &a.blue, BX0, BX1);
        for (int idx = threadIdx.x; idx < G * h; idx += kXT) {
            const int m = idx >> a.lgG, rr = idx & (G - 1);
            if (rr >= ns) continue;
            const float2 v = cur[idx];
            const float o0 = v.x * a.fct, o1 = v.y * a.fct;
            if (!a.along_cols) *reinterpret_cast<float2 *>(plane + (size_t)(s0 + rr) * a.img.pitch + 2 * m) = make_float2(o0, o1);
            else { plane[(size_t)(2 * m) * a.img.pitch + s0 + rr] = o0; plane[(size_t)(2 * m + 1) * a.img.pitch + s0 + rr] = o1; }
        }
    }
}

// one-row complex forward transform used once per plan to build bkf on the device with the same passes
__global__ void __launch_bounds__(kXT) k_blue_setup(XBlue b, const float *tab, const float2 *tbkf, float2 *out)
{
    extern __shared__ __align__(16) float xs[];
    float2 *X0 = reinterpret_cast<float2 *>(xs), *X1 = X0 + b.n2;
    Ctx c;
    c.r = 0; c.item0 = threadIdx.x; c.istride = kXT; c.GP = 1;
    for (int m = threadIdx.x; m < b.n2; m += kXT) X0[m] = tbkf[m];
    __syncthreads();
    float2 *cur = X0, *nxt = X1;
    cfft_all<true>(c, b, tab, cur, nxt, 1);
    for (int m = threadIdx.x; m < b.n2 / 2 + 1; m += kXT) out[m] = cur[m];
}

// ================================================================ host: plan + tables
// sincos_2pibyn<float>: exp(2 pi i k / n) from two double tables, product rounded to float
struct SinCos {
    size_t N, mask, shift;
    std::vector<double> v1, v2;
    static void calc(size_t x, size_t n, double ang, double *res)
    {
        x <<= 3;
        if (x < 4 * n) {
            if (x < 2 * n) {
                if (x < n) { res[0] = std::cos(double(x) * ang); res[1] = std::sin(double(x) * ang); return; }
                res[0] = std::sin(double(2 * n - x) * ang); res[1] = std::cos(double(2 * n - x) * ang); return;
            }
            x -= 2 * n;
            if (x < n) { res[0] = -std::sin(double(x) * ang); res[1] = std::cos(double(x) * ang); return; }
            res[0] = -std::cos(double(2 * n - x) * ang); res[1] = std::sin(double(2 * n - x) * ang); return;
        }
        x = 8 * n - x;
        if (x < 2 * n) {
            if (x < n) { res[0] = std::cos(double(x) * ang); res[1] = -std::sin(double(x) * ang); return; }
            res[0] = std::sin(double(2 * n - x) * ang); res[1] = -std::cos(double(2 * n - x) * ang); return;
        }
        x -= 2 * n;
        if (x < n) { res[0] = -std::sin(double(x) * ang); res[1] = -std::cos(double(x) * ang); return; }
        res[0] = -std::cos(double(2 * n - x) * ang); res[1] = -std::sin(double(2 * n - x) * ang);
    }
    explicit SinCos(size_t n) : N(n)
    {
        const long double pi = 3.141592653589793238462643383279502884197L;
        const double ang = double(0.25L * pi / (long double)n);
        const size_t nval = (n + 2) / 2;
        shift = 1;
        while ((size_t(1) << shift) * (size_t(1) << shift) < nval) ++shift;
        mask = (size_t(1) << shift) - 1;
        v1.resize(2 * (mask + 1));
        v1[0] = 1.0; v1[1] = 0.0;
        for (size_t i = 1; i < mask + 1; ++i) calc(i, n, ang, &v1[2 * i]);
        const size_t n2 = (nval + mask) / (mask + 1);
        v2.resize(2 * n2);
        v2[0] = 1.0; v2[1] = 0.0;
        for (size_t i = 1; i < n2; ++i) calc(i * (mask + 1), n, ang, &v2[2 * i]);
    }
    void get(size_t idx, float *re, float *im) const
    {
        const bool low = 2 * idx <= N;
        if (!low) idx = N - idx;
        const double *x1 = &v1[2 * (idx & mask)], *x2 = &v2[2 * (idx >> shift)];
        // volatile: keep the two products and the sum separately rounded (no host FMA contraction)
        volatile double rr = x1[0] * x2[0], ri = x1[1] * x2[1], ir = x1[0] * x2[1], ii = x1[1] * x2[0];
        *re = float(rr - ri);
        *im = low ? float(ir + ii) : -float(ir + ii);
    }
};

// fdiv magics of a complex sub-plan (see fdiv): valid while item * divisor < 2^32, true for every n2 a CTA can hold
void fill_magics(XBlue &b)
{
    auto magic = [](int d) -> unsigned { return d <= 1 ? 0u : (unsigned)((1ULL << 32) / (unsigned)d) + 1u; };
    int l1 = 1;
    for (int f = 0; f < b.nf; ++f) {
        const int ip = b.fct[f], ido = b.n2 / (l1 * ip);
        b.m_ido[f] = magic(ido);
        b.m_per[f] = magic(l1 * ido);
        l1 *= ip;
    }
    b.m_n2 = magic(b.n2);
}

size_t good_size_cmplx(size_t n)
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

size_t xfft_smem(int n, int G, const XBlue &b, int gt_max)
{
    const size_t GP = G;
    const size_t ab = (2 * (size_t)n * GP + 3) & ~(size_t)3;   // keeps the table 16-byte aligned
    return sizeof(float) * ab + sizeof(float2) * ((size_t)gt_max + 2 * (size_t)b.inst * b.n2 * GP);
}

}  // namespace

struct B2sXfftPlan {
    XArgs a;            // img, g, tab, along_cols, nseq, groups_per_plane are filled per launch
    XCArgs ca;          // complexify variant (even lengths > 1000)
    int cplx;
    size_t smem;
    float *d_tab;
};

namespace {
std::vector<int> prime_factors(int n)
{
    std::vector<int> f;
    for (int p = 2; (long long)p * p <= n; ++p)
        while (n % p == 0) { f.push_back(p); n /= p; }
    if (n > 1) f.push_back(n);
    return f;
}
// how scipy's (ducc0) float32 r2r transform of length n is evaluated on the rows of its 4-wide SIMD batches: 0 = real
// passes (rfftp, Bluestein passes for prime factors >= 135), 1 = half-length complex transform (rfftp_complexify, even
// lengths > 1000 whose half length has a prime factor >= 7), -1 = not mirrored (two Bluestein factors).
// Classification pinned empirically against scipy 1.18 for every even length in (1000, 3400): tests/test_oracle.py.
// The at most 3 rows of a sub-band that scipy processes outside its SIMD batches round differently when 8 divides the
// half length; they are evaluated like the others here.
int xfft_class(int n)
{
    if (n < 2) return -1;
    if (n <= 1000 || (n & 1)) return 0;
    const std::vector<int> f = prime_factors(n / 2);
    int big = 0;
    for (int p : f) big = std::max(big, p);
    if (big <= 5) return 0;                      // 5-smooth half length: plain real passes
    int n_blue = 0;
    for (int p : f) n_blue += p >= 110;          // prime factors >= 110 run as complex Bluestein passes (one supported)
    return n_blue <= 1 ? 1 : -1;
}
}  // namespace

int b2s_xfft_supported(int n) { return xfft_class(n) >= 0; }

static B2sXfftPlan *xfft_create_cplx(int n)
{
    B2sXfftPlan *pl = new B2sXfftPlan();
    pl->cplx = 1;
    XCArgs &a = pl->ca;
    memset(&a, 0, sizeof a);
    a.n = n;
    const int h = n / 2;
    XBlue &b = a.cp;
    b.n2 = h;
    std::vector<int> cf;
    int len = h;
    while ((len & 7) == 0) { cf.push_back(8); len >>= 3; }
    while ((len & 3) == 0) { cf.push_back(4); len >>= 2; }
    if ((len & 1) == 0) { len >>= 1; cf.push_back(2); std::swap(cf[0], cf.back()); }
    for (int d = 3; d * d <= len; d += 2)
        while ((len % d) == 0) { cf.push_back(d); len /= d; }
    if (len > 1) cf.push_back(len);
    if (cf.size() > 12) { delete pl; return nullptr; }
    b.nf = (int)cf.size();
    std::vector<float> tab;
    SinCos comp(h);
    int l1 = 1, blue_ip = 0, blue_need = 1;
    for (int k = 0; k < b.nf; ++k) {
        const int ip = cf[k], ido = h / (l1 * ip);
        b.fct[k] = ip;
        while (tab.size() & 1) tab.push_back(0.f);
        b.tw[k] = (int)tab.size();
        tab.resize(tab.size() + 2 * (size_t)(ip - 1) * (ido - 1), 0.f);
        for (int j = 1; j < ip; ++j)
            for (int i = 1; i < ido; ++i)
                comp.get((size_t)j * l1 * i, &tab[b.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1)],
                         &tab[b.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1) + 1]);
        if (ip >= 110) {
            blue_ip = ip;
            blue_need = l1 * ido;
        } else if (ip > 11) {
            a.cs[k] = (int)tab.size();
            tab.resize(tab.size() + 2 * (size_t)ip, 0.f);
            for (int j = 0; j < ip; ++j) comp.get((size_t)j * l1 * ido, &tab[a.cs[k] + 2 * j], &tab[a.cs[k] + 2 * j + 1]);
            a.wal_max = std::max(a.wal_max, ip);
        }
        l1 *= ip;
    }
    fill_magics(b);
    std::vector<float2> tbkf;
    XBlue &bl = a.blue;
    if (blue_ip) {   // Bluestein sub-plan: same construction as for the real passes (b2s_xfft_create)
        bl.ip = blue_ip;
        bl.n2 = (int)good_size_cmplx((size_t)blue_ip * 2 - 1);
        std::vector<int> bf;
        int ln = bl.n2;
        while ((ln & 7) == 0) { bf.push_back(8); ln >>= 3; }
        while ((ln & 3) == 0) { bf.push_back(4); ln >>= 2; }
        if ((ln & 1) == 0) { ln >>= 1; bf.push_back(2); std::swap(bf[0], bf.back()); }
        for (int d = 3; d * d <= ln; d += 2)
            while ((ln % d) == 0) { bf.push_back(d); ln /= d; }
        if (ln > 1) bf.push_back(ln);
        if (bf.size() > 12) { delete pl; return nullptr; }
        bl.nf = (int)bf.size();
        SinCos c2(bl.n2);
        int m1 = 1;
        for (int k = 0; k < bl.nf; ++k) {
            const int ip = bf[k], ido = bl.n2 / (m1 * ip);
            bl.fct[k] = ip;
            while (tab.size() & 1) tab.push_back(0.f);
            bl.tw[k] = (int)tab.size();
            tab.resize(tab.size() + 2 * (size_t)(ip - 1) * (ido - 1), 0.f);
            for (int j = 1; j < ip; ++j)
                for (int i = 1; i < ido; ++i)
                    c2.get((size_t)j * m1 * i, &tab[bl.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1)],
                           &tab[bl.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1) + 1]);
            m1 *= ip;
        }
        fill_magics(bl);
        while (tab.size() & 1) tab.push_back(0.f);
        bl.bk = (int)tab.size();
        tab.resize(tab.size() + 2 * (size_t)blue_ip, 0.f);
        SinCos tmp(2 * (size_t)blue_ip);
        tab[bl.bk] = 1.f; tab[bl.bk + 1] = 0.f;
        size_t coeff = 0;
        for (int m = 1; m < blue_ip; ++m) {
            coeff += 2 * (size_t)m - 1;
            if (coeff >= 2 * (size_t)blue_ip) coeff -= 2 * (size_t)blue_ip;
            tmp.get(coeff, &tab[bl.bk + 2 * m], &tab[bl.bk + 2 * m + 1]);
        }
        bl.bkf = (int)tab.size();
        tab.resize(tab.size() + 2 * (size_t)(bl.n2 / 2 + 1), 0.f);
        tbkf.assign(bl.n2, make_float2(0.f, 0.f));
        volatile float xn2 = 1.f / (float)bl.n2;
        for (int m = 0; m < blue_ip; ++m) {
            volatile float re = tab[bl.bk + 2 * m] * xn2, im = tab[bl.bk + 2 * m + 1] * xn2;
            tbkf[m] = make_float2(re, im);
            if (m) tbkf[bl.n2 - m] = tbkf[m];
        }
    }
    while (tab.size() & 1) tab.push_back(0.f);
    a.roots = (int)tab.size();
    tab.resize(tab.size() + 2 * (size_t)(h / 2 + 1), 0.f);
    SinCos rt(n);
    for (int i = 0; i <= h / 2; ++i) rt.get(i, &tab[a.roots + 2 * i], &tab[a.roots + 2 * i + 1]);
    a.fct = (float)(1.0L / (long double)n);
    auto smem_of = [&](int g) {
        return sizeof(float2) * (2 * (size_t)h * g + (size_t)a.wal_max + 2 + 2 * (size_t)bl.inst * bl.n2 * g);
    };
    int G = 0;
    for (int budget : {110 * 1024, 220 * 1024}) {
        for (int g = 16; g >= 1 && !G; g >>= 1) {
            if (blue_ip) {
                for (int inst = blue_need < 8 ? blue_need : 8; inst >= 1; --inst) {
                    bl.inst = inst;
                    if (smem_of(g) <= (size_t)budget) { G = g; break; }
                }
            } else if (smem_of(g) <= (size_t)budget) G = g;
        }
        if (G) break;
    }
    if (!G || cudaMalloc(&pl->d_tab, sizeof(float) * (tab.size() + 2)) != cudaSuccess) { delete pl; return nullptr; }
    cudaMemcpy(pl->d_tab, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice);
    if (blue_ip) {
        float2 *d_t = nullptr;
        cudaMalloc(&d_t, sizeof(float2) * bl.n2);
        cudaMemcpy(d_t, tbkf.data(), sizeof(float2) * bl.n2, cudaMemcpyHostToDevice);
        const size_t sm = sizeof(float2) * 4 * (size_t)bl.n2;
        cudaFuncSetAttribute(k_blue_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        k_blue_setup<<<1, kXT, sm>>>(bl, pl->d_tab, d_t, reinterpret_cast<float2 *>(pl->d_tab + bl.bkf));
        cudaDeviceSynchronize();
        cudaFree(d_t);
    }
    a.tab = pl->d_tab;
    a.G = G;
    while ((1 << a.lgG) < G) ++a.lgG;
    pl->smem = smem_of(G);
    return pl;
}

// builds the plan for length n (tables uploaded to the current device); returns nullptr when the length class is not
// covered or the work set does not fit in shared memory
B2sXfftPlan *b2s_xfft_create(int n)
{
    const int cls = xfft_class(n);
    if (cls < 0) return nullptr;
    if (cls == 1) return xfft_create_cplx(n);
    B2sXfftPlan *pl = new B2sXfftPlan();
    pl->cplx = 0;
    XArgs &a = pl->a;
    memset(&a, 0, sizeof a);
    a.n = n;
    // rfftp::factorize
    std::vector<int> fct;
    {
        int len = n;
        while ((len % 4) == 0) { fct.push_back(4); len >>= 2; }
        if ((len % 2) == 0) { len >>= 1; fct.push_back(2); std::swap(fct[0], fct.back()); }
        for (int d = 3; d * d <= len; d += 2)
            while ((len % d) == 0) { fct.push_back(d); len /= d; }
        if (len > 1) fct.push_back(len);
    }
    const int nf = (int)fct.size();
    if (nf > 12) { delete pl; return nullptr; }
    std::vector<float> tab;
    SinCos twid(n);
    std::vector<int> tw_off(nf), cs_off(nf, 0);
    int n_blue = 0, blue_ip = 0;
    {
        int l1 = 1;
        for (int k = 0; k < nf; ++k) {
            const int ip = fct[k], ido = n / (l1 * ip);
            while (tab.size() & 1) tab.push_back(0.f);
            tw_off[k] = (int)tab.size();
            tab.resize(tab.size() + (size_t)(ip - 1) * (ido - 1), 0.f);
            for (int j = 1; j < ip; ++j)
                for (int i = 1; i <= (ido - 1) / 2; ++i)
                    twid.get((size_t)j * l1 * i, &tab[tw_off[k] + (j - 1) * (ido - 1) + 2 * i - 2],
                             &tab[tw_off[k] + (j - 1) * (ido - 1) + 2 * i - 1]);
            if (ip > 5 && ip < 135) {
                // csarr (rfftp::comp_twiddle, "extra factors required by *g functions"), expanded to T[l-1][j-1]
                std::vector<float> tws(2 * (size_t)ip);
                tws[0] = 1.f; tws[1] = 0.f;
                for (int i = 2, ic = 2 * ip - 2; i <= ic; i += 2, ic -= 2) {
                    float re, im;
                    twid.get((size_t)(i / 2) * (n / ip), &re, &im);
                    tws[i] = re; tws[i + 1] = im; tws[ic] = re; tws[ic + 1] = -im;
                }
                const int ipph = (ip + 1) / 2, JP = kRow, rows = ipph - 1 + kLBMax;   // zero rows pad the last block of any block size
                while (tab.size() & 3) tab.push_back(0.f);
                cs_off[k] = (int)tab.size();
                tab.resize(tab.size() + 2 * (size_t)rows * JP, 0.f);
                for (int l = 1; l < ipph; ++l)
                    for (int j = 1; j < ipph; ++j) {
                        const int iang = (int)(((long long)l * j) % ip);
                        tab[cs_off[k] + 2 * ((size_t)(l - 1) * JP + (j - 1))] = tws[2 * iang];
                        tab[cs_off[k] + 2 * ((size_t)(l - 1) * JP + (j - 1)) + 1] = tws[2 * iang + 1];
                    }
            }
            if (ip >= 135) { ++n_blue; blue_ip = ip; }
            l1 *= ip;
        }
    }
    if (n_blue > 1) { delete pl; return nullptr; }
    a.nf = nf;
    a.negzero = -0.0f;
    a.gt_max = 0;
    bool magic_ok = true;
    auto magic = [&](int d, long long max_it) -> unsigned {
        if (d <= 1) return 0u;
        if (max_it * d >= (1LL << 32)) magic_ok = false;
        return (unsigned)((1ULL << 32) / (unsigned)d) + 1u;
    };
    auto make_pass = [&](int k, int l1, int ido) {
        const int ip = fct[k], ni = (ido - 1) / 2, ipph = (ip + 1) / 2;
        const long long max_it = (long long)n * 2 + 64;     // every item loop of a pass runs over fewer than 2n items
        XPass p{ip <= 5 ? ip : (ip < 135 ? 6 : 7), ip, l1, ido, tw_off[k], cs_off[k],
                magic(l1, max_it), magic(ni, max_it), magic(l1 * ni, max_it), magic(ido, max_it), magic(ido * l1, max_it), kLBMax};
        if (p.kind == 6) a.gt_max = std::max(a.gt_max, (ipph - 1 + kLBMax) * kRow);
        return p;
    };
    {
        int l1 = n;
        for (int k1 = 0; k1 < nf; ++k1) {   // forward: last factor first
            const int k = nf - 1 - k1, ip = fct[k], ido = n / l1;
            l1 /= ip;
            a.fwd[k1] = make_pass(k, l1, ido);
        }
        l1 = 1;
        for (int k = 0; k < nf; ++k) {
            const int ip = fct[k], ido = n / (ip * l1);
            a.bwd[k] = make_pass(k, l1, ido);
            l1 *= ip;
        }
    }
    if (!magic_ok) { delete pl; return nullptr; }
    XBlue &b = a.blue;
    std::vector<float2> tbkf;
    if (n_blue) {
        b.ip = blue_ip;
        b.n2 = (int)good_size_cmplx((size_t)blue_ip * 2 - 1);
        // cfftp::factorize
        std::vector<int> cf;
        int len = b.n2;
        while ((len & 7) == 0) { cf.push_back(8); len >>= 3; }
        while ((len & 3) == 0) { cf.push_back(4); len >>= 2; }
        if ((len & 1) == 0) { len >>= 1; cf.push_back(2); std::swap(cf[0], cf.back()); }
        for (int d = 3; d * d <= len; d += 2)
            while ((len % d) == 0) { cf.push_back(d); len /= d; }
        if (len > 1) cf.push_back(len);
        if (cf.size() > 12) { delete pl; return nullptr; }
        b.nf = (int)cf.size();
        SinCos comp(b.n2);
        int l1 = 1;
        for (int k = 0; k < b.nf; ++k) {
            const int ip = cf[k], ido = b.n2 / (l1 * ip);
            b.fct[k] = ip;
            while (tab.size() & 1) tab.push_back(0.f);
            b.tw[k] = (int)tab.size();
            tab.resize(tab.size() + 2 * (size_t)(ip - 1) * (ido - 1), 0.f);
            for (int j = 1; j < ip; ++j)
                for (int i = 1; i < ido; ++i)
                    comp.get((size_t)j * l1 * i, &tab[b.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1)],
                             &tab[b.tw[k] + 2 * ((j - 1) * (ido - 1) + i - 1) + 1]);
            l1 *= ip;
        }
        fill_magics(b);
        // bk
        while (tab.size() & 1) tab.push_back(0.f);
        b.bk = (int)tab.size();
        tab.resize(tab.size() + 2 * (size_t)blue_ip, 0.f);
        SinCos tmp(2 * (size_t)blue_ip);
        tab[b.bk] = 1.f; tab[b.bk + 1] = 0.f;
        size_t coeff = 0;
        for (int m = 1; m < blue_ip; ++m) {
            coeff += 2 * (size_t)m - 1;
            if (coeff >= 2 * (size_t)blue_ip) coeff -= 2 * (size_t)blue_ip;
            tmp.get(coeff, &tab[b.bk + 2 * m], &tab[b.bk + 2 * m + 1]);
        }
        b.bkf = (int)tab.size();
        tab.resize(tab.size() + 2 * (size_t)(b.n2 / 2 + 1), 0.f);
        // zero-padded, normalised b_k whose transform the device computes below
        tbkf.assign(b.n2, make_float2(0.f, 0.f));
        volatile float xn2 = 1.f / (float)b.n2;
        for (int m = 0; m < blue_ip; ++m) {
            volatile float re = tab[b.bk + 2 * m] * xn2, im = tab[b.bk + 2 * m + 1] * xn2;
            tbkf[m] = make_float2(re, im);
            if (m) tbkf[b.n2 - m] = tbkf[m];
        }
    }
    a.fct = (float)(1.0L / (long double)n);
    if (cudaMalloc(&pl->d_tab, sizeof(float) * (tab.size() + 2)) != cudaSuccess) { delete pl; return nullptr; }
    cudaMemcpy(pl->d_tab, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice);
    a.tab = pl->d_tab;
    if (n_blue) {
        float2 *d_t = nullptr;
        cudaMalloc(&d_t, sizeof(float2) * b.n2);
        cudaMemcpy(d_t, tbkf.data(), sizeof(float2) * b.n2, cudaMemcpyHostToDevice);
        const size_t sm = sizeof(float2) * 4 * (size_t)b.n2;
        cudaFuncSetAttribute(k_blue_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        k_blue_setup<<<1, kXT, sm>>>(b, pl->d_tab, d_t, reinterpret_cast<float2 *>(pl->d_tab + b.bkf));
        cudaDeviceSynchronize();
        cudaFree(d_t);
    }
    // rows per CTA: the largest power of two whose work set allows two CTAs per SM, else one
    int G = 0;
    for (int budget : {110 * 1024, 220 * 1024}) {
        for (int g = 16; g >= 1 && !G; g >>= 1) {
            if (n_blue) {
                // concurrent Bluestein instances: as many as fit, at most what one pass needs
                int need = 1;
                for (int f = 0; f < nf; ++f)
                    if (a.fwd[f].kind == 7) need = a.fwd[f].l1 * (1 + (a.fwd[f].ido - 1) / 2);
                for (int inst = need < 8 ? need : 8; inst >= 1; --inst) {
                    b.inst = inst;
                    if (xfft_smem(n, g, b, a.gt_max) <= (size_t)budget) { G = g; break; }
                }
            } else if (xfft_smem(n, g, b, a.gt_max) <= (size_t)budget) G = g;
        }
        if (G) break;
    }
    if (n_blue) {   // experiments: B2S_XFFT_G_BLUE forces the rows per CTA of a Bluestein plan when its work set fits
        static const int forced = getenv("B2S_XFFT_G_BLUE") ? atoi(getenv("B2S_XFFT_G_BLUE")) : 0;
        if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) {
            const int keep = b.inst;
            b.inst = 1;
            if (xfft_smem(n, forced, b, a.gt_max) <= (size_t)220 * 1024) G = forced; else b.inst = keep;
        }
    }
    if (!G) { cudaFree(pl->d_tab); delete pl; return nullptr; }
    a.G = G;
    a.lgG = 0;
    while ((1 << a.lgG) < G) ++a.lgG;
    pl->smem = xfft_smem(n, G, b, a.gt_max);
    // block size of the O(radix^2) phase (outputs per thread): 8.  Smaller blocks fill the rounds of the CTA's item slots
    // better (n = 1333 = 31 * 43, 64 slots: LB = 8 -> 93 and 86 items, 73 % / 67 % of two rounds; LB = 6 / 4 -> 97 % / 90 %) but
    // measured no faster — the kernel is issue-bound, a warp idling at a barrier costs nothing while others issue
    // (profiles/r02_notch_experiments.md).  B2S_XFFT_LB = 4 .. 8 forces a size for experiments.
    {
        static const int forced = getenv("B2S_XFFT_LB") ? atoi(getenv("B2S_XFFT_LB")) : 0;
        const int lb = (forced >= 4 && forced <= kLBMax) ? forced : kLBMax;
        for (int f = 0; f < nf; ++f) { a.fwd[f].lb = lb; a.bwd[f].lb = lb; }
    }
    return pl;
}

void b2s_xfft_destroy(B2sXfftPlan *pl)
{
    if (!pl) return;
    cudaFree(pl->d_tab);
    delete pl;
}

// CTAs per SM the grid is capped at.  Measured on the bench workload (32 planes, 167 groups of 8 rows per plane):
// 8 -> 31.7 us/plane at level 1 (CTAs loop over 4 or 5 groups: a 10 % tail), 40 (one group per CTA) -> 29.5
static int xfft_cap_mult()
{
    static const int m = getenv("B2S_XFFT_CAP") ? atoi(getenv("B2S_XFFT_CAP")) : 40;
    return m > 0 ? m : 1;
}

void b2s_launch_notch_exact(const B2sXfftPlan *pl, const float *d_notch, const B2sImg &img, int along_cols, int n_planes,
                            int sm_count, cudaStream_t s)
{
    if (pl->cplx) {
        XCArgs a = pl->ca;
        a.img = img;
        a.g = d_notch;
        a.along_cols = along_cols;
        a.nseq = along_cols ? img.cols : img.rows;
        a.groups_per_plane = (a.nseq + a.G - 1) / a.G;
        cudaFuncSetAttribute(k_notch_cplx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem);
        int bx = a.groups_per_plane;
        const int cap = (sm_count * xfft_cap_mult() + n_planes - 1) / n_planes;
        if (bx > cap) bx = cap > 0 ? cap : 1;
        k_notch_cplx<<<dim3(bx, n_planes), kXT, pl->smem, s>>>(a);
        return;
    }
    XArgs a = pl->a;
    a.img = img;
    a.g = d_notch;
    a.along_cols = along_cols;
    a.nseq = along_cols ? img.cols : img.rows;
    a.groups_per_plane = (a.nseq + a.G - 1) / a.G;
    cudaFuncSetAttribute(k_notch_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem);
    int bx = a.groups_per_plane;
    const int cap = (sm_count * xfft_cap_mult() + n_planes - 1) / n_planes;
    if (bx > cap) bx = cap > 0 ? cap : 1;
    k_notch_exact<<<dim3(bx, n_planes), kXT, pl->smem, s>>>(a);
}
