// Bit-exact float32 log1p / expm1 for the device.
//
// The reference evaluates log1p/expm1 in float32 through numexpr (pystripe/core.py:184,194), i.e. through the
// host libm.  To be bit-identical the device cannot use CUDA's log1pf/expm1f (different polynomials); these two
// functions follow the classic fdlibm float algorithms (the ones glibc <= 2.40 ships as log1pf / expm1f):
// same argument reduction, same constants, same operation order, every operation a single IEEE-754 binary32
// operation (compile with -fmad=false; on the host -ffp-contract=off).
// Verified against the host libm over every finite float (tools/check_libm_mirror.c, tests/test_libm_mirror.py).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD static inline
#endif

B2S_HD int32_t b2s_f2i(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4); return i;
#endif
}
B2S_HD float b2s_i2f(int32_t i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

B2S_HD float b2s_log1pf(float x)
{
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    const float Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f,
                Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f,
                Lp7 = 1.4798198640e-01f;
    float hfsq, f = 0.f, c = 0.f, s, z, R, u;
    int32_t k, hx, hu = 0, ax;

    hx = b2s_f2i(x);
    ax = hx & 0x7fffffff;
    k = 1;
    if (hx < 0x3ed413d7) {                       /* x < 0.41422 */
        if (ax >= 0x3f800000) {                  /* x <= -1 */
            if (x == -1.0f) return -b2s_i2f(0x7f800000);
            return b2s_i2f(0x7fc00000);
        }
        if (ax < 0x31000000) {                   /* |x| < 2^-29 */
            if (ax < 0x24800000) return x;
            return x - x * x * 0.5f;
        }
        if (hx > 0 || hx <= (int32_t)0xbe95f61f) { k = 0; f = x; hu = 1; }
    }
    if (hx >= 0x7f800000) return x + x;
    if (k != 0) {
        if (hx < 0x5a000000) {
            u = 1.0f + x;
            hu = b2s_f2i(u);
            k = (hu >> 23) - 127;
            c = (k > 0) ? 1.0f - (u - x) : x - (u - 1.0f);
            c /= u;
        } else {
            u = x;
            hu = b2s_f2i(u);
            k = (hu >> 23) - 127;
            c = 0.f;
        }
        hu &= 0x007fffff;
        if (hu < 0x3504f7) {
            u = b2s_i2f(hu | 0x3f800000);
        } else {
            k += 1;
            u = b2s_i2f(hu | 0x3f000000);
            hu = (0x00800000 - hu) >> 2;
        }
        f = u - 1.0f;
    }
    hfsq = 0.5f * f * f;
    if (hu == 0) {                               /* |f| < 2^-20 */
        if (f == 0.f) {
            if (k == 0) return 0.f;
            c += k * ln2_lo;
            return k * ln2_hi + c;
        }
        R = hfsq * (1.0f - 0.66666666666666666f * f);
        if (k == 0) return f - R;
        return k * ln2_hi - ((R - (k * ln2_lo + c)) - f);
    }
    s = f / (2.0f + f);
    z = s * s;
    R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return k * ln2_hi - ((hfsq - (s * (hfsq + R) + (k * ln2_lo + c))) - f);
}

B2S_HD float b2s_expm1f(float x)
{
    const float o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                invln2 = 1.4426950216e+00f;
    const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f,
                Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    const float huge = 1.0e+30f, tiny = 1.0e-30f;
    float y, hi, lo, c = 0.f, t, e, hxs, hfx, r1;
    int32_t k, xsb;
    uint32_t hx;

    hx = (uint32_t)b2s_f2i(x);
    xsb = (int32_t)(hx & 0x80000000u);
    hx &= 0x7fffffffu;

    if (hx >= 0x4195b844u) {                     /* |x| >= 27 ln2 */
        if (hx >= 0x42b17218u) {                 /* |x| >= 88.72 */
            if (hx > 0x7f800000u) return x + x;
            if (hx == 0x7f800000u) return (xsb == 0) ? x : -1.0f;
            if (x > o_threshold) return huge * huge;
        }
        if (xsb != 0) return tiny - 1.0f;
    }
    if (hx > 0x3eb17218u) {                      /* |x| > 0.5 ln2 */
        if (hx < 0x3F851592u) {                  /* |x| < 1.5 ln2 */
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else          { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int32_t)(invln2 * x + ((xsb == 0) ? 0.5f : -0.5f));
            t = (float)k;
            hi = x - t * ln2_hi;
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x33000000u) {               /* |x| < 2^-25 */
        t = huge + x;
        return x - (t - (huge + x));
    } else {
        k = 0;
    }
    hfx = 0.5f * x;
    hxs = x * hfx;
    r1 = 1.0f + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0f - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0f - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5f * (x - e) - 0.5f;
    if (k == 1) {
        if (x < -0.25f) return -2.0f * (e - (x + 0.5f));
        return 1.0f + 2.0f * (x - e);
    }
    if (k <= -2 || k > 56) {
        y = 1.0f - (e - x);
        y = b2s_i2f(b2s_f2i(y) + (k << 23));
        return y - 1.0f;
    }
    t = 1.0f;
    if (k < 23) {
        t = b2s_i2f(0x3f800000 - (0x1000000 >> k));
        y = t - (e - x);
        y = b2s_i2f(b2s_f2i(y) + (k << 23));
    } else {
        t = b2s_i2f((0x7f - k) << 23);
        y = x - (e + t);
        y += 1.0f;
        y = b2s_i2f(b2s_f2i(y) + (k << 23));
    }
    return y;
}


// ---- device hot paths ---------------------------------------------------------------------------------------------
// The same operations as above, in the same order, for the argument range pixel data lives in; everything else falls
// back to the general routine.  (One range test replaces the chain of special-case branches the general routines walk
// for every pixel; tests/test_gpu_parity.py::test_device_math_is_bit_exact covers both paths against the host libm.)
#if defined(__CUDACC__)
// Correctly rounded a / b without the IEEE-division subroutine, for finite operands whose quotient is zero or a normal
// number far from overflow (true for every division in the two hot paths: |quotient| in [2^-60, 2^20] or exactly 0):
// reciprocal estimate, one Newton step, quotient, two exact-residual corrections (Markstein).  The residual a - b*q is
// exactly representable once q is within an ulp, so the last fma rounds the true quotient once.
__device__ __forceinline__ float b2s_div_hot(float a, float b)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
    float q = __fmul_rn(a, r);
    q = fmaf(fmaf(-b, q, a), r, q);
    q = fmaf(fmaf(-b, q, a), r, q);
    return q;
}

__device__ __forceinline__ float b2s_expm1f_dev(float x)
{
    if (!(x >= 1.1f && x < 38.0f)) return b2s_expm1f(x);      // here 2 <= k <= 55
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f;
    const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f,
                Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    const int32_t k = (int32_t)(invln2 * x + 0.5f);
    const float t = (float)k;
    const float hi = x - t * ln2_hi, lo = t * ln2_lo;
    const float xr = hi - lo;
    const float c = (hi - xr) - lo;
    const float hfx = 0.5f * xr;
    const float hxs = xr * hfx;
    const float r1 = 1.0f + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    const float tt = 3.0f - r1 * hfx;
    float e = hxs * b2s_div_hot(r1 - tt, 6.0f - xr * tt);
    e = (xr * (e - c) - c);
    e -= hxs;
    float y;
    if (k < 23) {
        y = b2s_i2f(0x3f800000 - (0x1000000 >> k)) - (e - xr);
    } else {
        y = xr - (e + b2s_i2f((0x7f - k) << 23));
        y += 1.0f;
    }
    return b2s_i2f(b2s_f2i(y) + (k << 23));
}

// Branch-free variants of the two hot paths for loops that evaluate several pixels per thread: the same operations in
// the same order, but an argument outside the hot range (or log1p's |f| < 2^-20 case) only raises `bad` — the caller
// re-evaluates such a group of pixels with b2s_log1pf_dev / b2s_expm1f_dev.  One reconvergence point per group instead of
// three per pixel (BRA + BSSY + BSYNC were 14 % of k_prologue's instructions, profiles/r02_pre_tma_pro.source.csv.gz).
__device__ __forceinline__ float b2s_log1pf_hot(float x, bool &bad)
{
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    const float Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f,
                Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f,
                Lp7 = 1.4798198640e-01f;
    bad |= !(x >= 0.5f && x < 1.0e9f);
    float u = 1.0f + x;
    int32_t hu = b2s_f2i(u);
    int32_t k = (hu >> 23) - 127;
    float c = (k > 0) ? 1.0f - (u - x) : x - (u - 1.0f);
    c = b2s_div_hot(c, u);
    hu &= 0x007fffff;
    const bool low = hu < 0x3504f7;
    k += low ? 0 : 1;
    u = b2s_i2f(hu | (low ? 0x3f800000 : 0x3f000000));
    hu = low ? hu : ((0x00800000 - hu) >> 2);
    bad |= hu == 0;
    const float f = u - 1.0f;
    const float hfsq = 0.5f * f * f;
    const float s = b2s_div_hot(f, 2.0f + f);
    const float z = s * s;
    const float R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    const float kf = (float)k;
    return kf * ln2_hi - ((hfsq - (s * (hfsq + R) + (kf * ln2_lo + c))) - f);
}

__device__ __forceinline__ float b2s_expm1f_hot(float x, bool &bad)
{
    bad |= !(x >= 1.1f && x < 15.5f);                           // here 2 <= k <= 22: one form of the final scaling
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f;
    const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f,
                Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    const int32_t k = (int32_t)(invln2 * x + 0.5f);
    const float t = (float)k;
    const float hi = x - t * ln2_hi, lo = t * ln2_lo;
    const float xr = hi - lo;
    const float c = (hi - xr) - lo;
    const float hfx = 0.5f * xr;
    const float hxs = xr * hfx;
    const float r1 = 1.0f + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    const float tt = 3.0f - r1 * hfx;
    float e = hxs * b2s_div_hot(r1 - tt, 6.0f - xr * tt);
    e = (xr * (e - c) - c);
    e -= hxs;
    const float y = b2s_i2f(0x3f800000 - (0x1000000 >> (k & 31))) - (e - xr);
    return b2s_i2f(b2s_f2i(y) + (k << 23));
}

__device__ __forceinline__ float b2s_log1pf_dev(float x)
{
    if (!(x >= 0.5f && x < 1.0e9f)) return b2s_log1pf(x);     // here k != 0 and hx < 0x5a000000
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
    const float Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f,
                Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f,
                Lp7 = 1.4798198640e-01f;
    float u = 1.0f + x;
    int32_t hu = b2s_f2i(u);
    int32_t k = (hu >> 23) - 127;
    float c = (k > 0) ? 1.0f - (u - x) : x - (u - 1.0f);
    c = b2s_div_hot(c, u);
    hu &= 0x007fffff;
    if (hu < 0x3504f7) {
        u = b2s_i2f(hu | 0x3f800000);
    } else {
        k += 1;
        u = b2s_i2f(hu | 0x3f000000);
        hu = (0x00800000 - hu) >> 2;
    }
    if (hu == 0) return b2s_log1pf(x);                          // |f| < 2^-20: rare, general routine
    const float f = u - 1.0f;
    const float hfsq = 0.5f * f * f;
    const float s = b2s_div_hot(f, 2.0f + f);
    const float z = s * s;
    const float R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    const float kf = (float)k;
    return kf * ln2_hi - ((hfsq - (s * (hfsq + R) + (kf * ln2_lo + c))) - f);
}
#endif
