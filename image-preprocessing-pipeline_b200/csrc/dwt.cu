// Separable 2-channel orthogonal filter banks (analysis + synthesis) for one decomposition level, batched over planes.
//
// Replaces the PyWavelets C loops the reference reaches through pywt.wavedec2 / waverec2
// (pystripe/core.py:927, :939): float_downsampling_convolution (MODE_SYMMETRIC) and
// float_upsampling_convolution_valid_sf.  Arithmetic contract in `exact` mode (default): float32, every tap is one
// multiply followed by one add into a running sum that starts at 0, taps visited in the reference order
// (analysis: j = 0..F-1, except right-overhang outputs which visit the reflected taps first; synthesis: j = 0..F/2-1
// for the low-pass branch, then the high-pass branch, then one add).  axis -2 (rows index) is analysed first and
// synthesised last, as pywt's dwtn / idwtn do.
//
// One CTA computes a TY x TX tile of all four sub-bands: the (2TY+F-2) x (2TX+F-2) input window is staged in shared
// memory with the half-sample symmetric extension resolved at load time, the axis -2 pass writes its two
// half-height images to a second shared buffer, the axis -1 pass reads them with conflict-free 128-bit shared loads
// and stores 128-bit coalesced rows.  No intermediate ever goes to HBM.
#include "b2s_internal.h"

namespace {

constexpr int kNT = 256;
// forward tile: TY x TX outputs per sub-band, each thread RY (axis -2 pass) / RX (axis -1 pass) consecutive outputs
constexpr int kTY = 32, kTX = 64, kRY = 4, kRX = 4;
// inverse tile: TQ x TP coefficient positions -> 2TQ x 2TP outputs
constexpr int kTQ = 32, kTP = 64;

__host__ __device__ inline int pitch_mod8_is4(int n)  // >= n, multiple of 4, (p/4) odd
{
    int p = (n + 3) & ~3;
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}
__host__ __device__ inline int pitch_mod32_is16(int n)  // >= n, multiple of 4, (p/4) % 8 == 4
{
    int p = (n + 3) & ~3;
    while (((p >> 2) & 7) != 4) p += 4;
    return p;
}

struct FwdGeom {
    int rin_y, rin_x, pitch;
    __host__ __device__ explicit FwdGeom(int F)
    {
        rin_y = 2 * kTY + F - 2;
        rin_x = 2 * kTX + F - 2;
        pitch = pitch_mod8_is4(rin_x + 2);
    }
    __host__ __device__ size_t smem_floats() const { return (size_t)(rin_y + 2 * kTY) * pitch + 8; }
};

struct InvGeom {
    int H, rq, rp, ps, pm;
    __host__ __device__ explicit InvGeom(int F)
    {
        H = F / 2;
        rq = kTQ + H - 1;
        rp = kTP + H - 1;
        ps = pitch_mod32_is16(rp + 3);
        pm = pitch_mod8_is4(2 * kTP);
    }
    __host__ __device__ size_t smem_floats() const { return (size_t)4 * rq * ps + (size_t)2 * rq * pm + 8; }
};

__device__ __forceinline__ int sym_ext(int i, int n)
{
    const int p = 2 * n;
    int t = i % p;
    if (t < 0) t += p;
    return t < n ? t : p - 1 - t;
}

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float a, float b)
{
    if (EXACT) return __fadd_rn(acc, __fmul_rn(a, b));  // never contracted into an FMA
    return fmaf(a, b, acc);
}

// analysis output whose window overhangs the right edge (i = 2o+1 >= n): pywt visits the reflected part with the
// filter index descending from i-n to 0, then the in-range part ascending.  `w` points at the extended sample that
// pairs with tap 0 (stride between samples = step, towards lower indices for higher taps).
template <bool EXACT>
__device__ __forceinline__ float analysis_right_edge(const float *filt, int F, const float *w0, int step, int over)
{
    float s = 0.f;
    for (int t = 0; t < F; ++t) {
        const int j = (t <= over) ? (over - t) : t;
        s = mac<EXACT>(s, filt[j], w0[-j * step]);
    }
    return s;
}

// ------------------------------------------------------------------------------------------------ forward
template <int F_, bool EXACT>
__global__ void __launch_bounds__(kNT, 2)
k_dwt_fwd(const B2sTaps taps, const B2sImg in, const B2sImg cA, const B2sImg cH, const B2sImg cV, const B2sImg cD)
{
    extern __shared__ __align__(16) float smem[];
    const int F = F_ > 0 ? F_ : taps.F;
    const FwdGeom g(F);
    const int P = g.pitch;
    float *s_in = smem;
    float *s_mid = smem + (size_t)g.rin_y * P;

    const int tid = threadIdx.x;
    const int oy0 = blockIdx.y * kTY, ox0 = blockIdx.x * kTX;
    const int ny = in.rows, nx = in.cols;
    const int my = cA.rows, mx = cA.cols;
    const float *src = in.ptr + (size_t)blockIdx.z * in.plane_stride;
    const int gy0 = 2 * oy0 - F + 2, gx0 = 2 * ox0 - F + 2;

    // ---- stage the extended input window
    {
        const bool x_inside = gx0 >= 0 && gx0 + g.rin_x <= nx;
        const int warp = tid >> 5, lane = tid & 31;
        for (int ry = warp; ry < g.rin_y; ry += kNT / 32) {
            const int sy = sym_ext(gy0 + ry, ny);
            const float *srow = src + (size_t)sy * in.pitch;
            float *drow = s_in + ry * P;
            if (x_inside) {
                for (int rx = lane; rx < g.rin_x; rx += 32) drow[rx] = __ldg(srow + gx0 + rx);
            } else {
                for (int rx = lane; rx < g.rin_x; rx += 32) drow[rx] = __ldg(srow + sym_ext(gx0 + rx, nx));
            }
        }
    }
    __syncthreads();

    // ---- axis -2 (rows index) pass: s_in -> s_mid[0..TY) = low-pass rows, s_mid[TY..2TY) = high-pass rows
    {
        constexpr int NGY = kTY / kRY;
        const int items = NGY * g.rin_x;
        for (int it = tid; it < items; it += kNT) {
            const int gy = it / g.rin_x;
            const int rx = it - gy * g.rin_x;
            const float *col = s_in + (2 * kRY * gy) * P + rx;
            float lo[kRY], hi[kRY];
            if (F_ > 0) {
                constexpr int FW = F_ > 0 ? F_ : 2;
                float w[2 * kRY + FW - 2];
#pragma unroll
                for (int t = 0; t < 2 * kRY + FW - 2; ++t) w[t] = col[t * P];
#pragma unroll
                for (int r = 0; r < kRY; ++r) {
                    float a = 0.f, d = 0.f;
#pragma unroll
                    for (int j = 0; j < FW; ++j) {
                        const float v = w[2 * r + FW - 1 - j];
                        a = mac<EXACT>(a, taps.dec_lo[j], v);
                        d = mac<EXACT>(d, taps.dec_hi[j], v);
                    }
                    lo[r] = a;
                    hi[r] = d;
                }
            } else {
#pragma unroll
                for (int r = 0; r < kRY; ++r) {
                    float a = 0.f, d = 0.f;
                    const float *w0 = col + (2 * r + F - 1) * P;
                    for (int j = 0; j < F; ++j) {
                        const float v = w0[-j * P];
                        a = mac<EXACT>(a, taps.dec_lo[j], v);
                        d = mac<EXACT>(d, taps.dec_hi[j], v);
                    }
                    lo[r] = a;
                    hi[r] = d;
                }
            }
            if (EXACT) {
#pragma unroll
                for (int r = 0; r < kRY; ++r) {
                    const int i = 2 * (oy0 + kRY * gy + r) + 1;
                    if (i >= ny && i - ny <= F - 2) {
                        const float *w0 = col + (2 * r + F - 1) * P;
                        lo[r] = analysis_right_edge<true>(taps.dec_lo, F, w0, P, i - ny);
                        hi[r] = analysis_right_edge<true>(taps.dec_hi, F, w0, P, i - ny);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kRY; ++r) {
                s_mid[(kRY * gy + r) * P + rx] = lo[r];
                s_mid[(kTY + kRY * gy + r) * P + rx] = hi[r];
            }
        }
    }
    __syncthreads();

    // ---- axis -1 pass: each lane owns RX consecutive output columns of one s_mid row
    {
        const int warp = tid >> 5, lane = tid & 31;
        const int gxl = lane & 3, rsub = lane >> 2;
        constexpr int NGX = kTX / kRX;      // 16 column groups
        constexpr int N_GB = NGX / 4;       // 4 group blocks
        constexpr int N_RB = (2 * kTY) / 8; // 8 row blocks
        float *pA = cA.ptr + (size_t)blockIdx.z * cA.plane_stride;
        float *pH = cH.ptr + (size_t)blockIdx.z * cH.plane_stride;
        float *pV = cV.ptr + (size_t)blockIdx.z * cV.plane_stride;
        float *pD = cD.ptr + (size_t)blockIdx.z * cD.plane_stride;
        for (int wi = warp; wi < N_RB * N_GB; wi += kNT / 32) {
            const int rb = wi / N_GB, gb = wi - rb * N_GB;
            const int rm = rb * 8 + rsub;
            const int gx = gb * 4 + gxl;
            const float *row = s_mid + rm * P + 2 * kRX * gx;
            float a[kRX], d[kRX];
            if (F_ > 0) {
                constexpr int FW = F_ > 0 ? F_ : 2;
                constexpr int NW = (2 * kRX + FW - 2 + 3) / 4;
                float w[NW * 4];
#pragma unroll
                for (int t = 0; t < NW; ++t) {
                    const float4 v = *reinterpret_cast<const float4 *>(row + 4 * t);
                    w[4 * t] = v.x; w[4 * t + 1] = v.y; w[4 * t + 2] = v.z; w[4 * t + 3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < kRX; ++c) {
                    float sa = 0.f, sd = 0.f;
#pragma unroll
                    for (int j = 0; j < FW; ++j) {
                        const float v = w[2 * c + FW - 1 - j];
                        sa = mac<EXACT>(sa, taps.dec_lo[j], v);
                        sd = mac<EXACT>(sd, taps.dec_hi[j], v);
                    }
                    a[c] = sa;
                    d[c] = sd;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kRX; ++c) {
                    float sa = 0.f, sd = 0.f;
                    const float *w0 = row + 2 * c + F - 1;
                    for (int j = 0; j < F; ++j) {
                        const float v = w0[-j];
                        sa = mac<EXACT>(sa, taps.dec_lo[j], v);
                        sd = mac<EXACT>(sd, taps.dec_hi[j], v);
                    }
                    a[c] = sa;
                    d[c] = sd;
                }
            }
            const int oxb = ox0 + kRX * gx;
            if (EXACT) {
#pragma unroll
                for (int c = 0; c < kRX; ++c) {
                    const int i = 2 * (oxb + c) + 1;
                    if (i >= nx && i - nx <= F - 2) {
                        const float *w0 = row + 2 * c + F - 1;
                        a[c] = analysis_right_edge<true>(taps.dec_lo, F, w0, 1, i - nx);
                        d[c] = analysis_right_edge<true>(taps.dec_hi, F, w0, 1, i - nx);
                    }
                }
            }
            const bool low_rows = rm < kTY;
            const int oy = oy0 + (low_rows ? rm : rm - kTY);
            if (oy < my && oxb < mx) {
                float *da = (low_rows ? pA : pH) + (size_t)oy * cA.pitch + oxb;  // all four share pitch
                float *dd = (low_rows ? pV : pD) + (size_t)oy * cA.pitch + oxb;
                if (oxb + kRX <= mx) {
                    *reinterpret_cast<float4 *>(da) = make_float4(a[0], a[1], a[2], a[3]);
                    *reinterpret_cast<float4 *>(dd) = make_float4(d[0], d[1], d[2], d[3]);
                } else {
#pragma unroll
                    for (int c = 0; c < kRX; ++c)
                        if (oxb + c < mx) { da[c] = a[c]; dd[c] = d[c]; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ inverse
template <int F_, bool EXACT>
__global__ void __launch_bounds__(kNT, 2)
k_dwt_inv(const B2sTaps taps, const B2sImg cA, const B2sImg cH, const B2sImg cV, const B2sImg cD, const B2sImg out)
{
    extern __shared__ __align__(16) float smem[];
    const int F = F_ > 0 ? F_ : taps.F;
    const InvGeom g(F);
    const int H = g.H, PS = g.ps, PM = g.pm, RQ = g.rq, RP = g.rp;
    float *s_sub = smem;                         // [4][RQ][PS] : 0 = cA, 1 = cV, 2 = cH, 3 = cD
    float *s_mid = smem + (size_t)4 * RQ * PS;   // [2][RQ][PM] : 0 = a (from cA,cV), 1 = d (from cH,cD)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.y * kTQ, p0 = blockIdx.x * kTP;
    const int my = cH.rows, mx = cH.cols;

    // ---- stage the four sub-band windows (zero beyond the sub-band: only feeds outputs that are not stored)
    {
        const float *base[4] = {cA.ptr + (size_t)blockIdx.z * cA.plane_stride, cV.ptr + (size_t)blockIdx.z * cV.plane_stride,
                                cH.ptr + (size_t)blockIdx.z * cH.plane_stride, cD.ptr + (size_t)blockIdx.z * cD.plane_stride};
        const int pitch[4] = {cA.pitch, cV.pitch, cH.pitch, cD.pitch};
        for (int r = warp; r < 4 * RQ; r += kNT / 32) {
            const int sb = r / RQ, ry = r - sb * RQ;
            const int y = q0 + ry;
            float *drow = s_sub + (size_t)r * PS;
            if (y < my) {
                const float *srow = base[sb] + (size_t)y * pitch[sb] + p0;
                for (int rx = lane; rx < PS; rx += 32) drow[rx] = (rx < RP && p0 + rx < mx) ? __ldg(srow + rx) : 0.f;
            } else {
                for (int rx = lane; rx < PS; rx += 32) drow[rx] = 0.f;
            }
        }
    }
    __syncthreads();

    // ---- axis -1 synthesis: (cA,cV) -> a, (cH,cD) -> d   [idwtn handles the last axis first]
    {
        const int gxl = lane & 3, rsub = lane >> 2;
        constexpr int N_GB = (kTP / kRX) / 4;  // 4
        const int n_rb = (2 * RQ + 7) / 8;
        for (int wi = warp; wi < n_rb * N_GB; wi += kNT / 32) {
            const int rb = wi / N_GB, gb = wi - rb * N_GB;
            const int rr = rb * 8 + rsub;
            if (rr >= 2 * RQ) continue;
            const int sel = rr >= RQ ? 1 : 0;
            const int ry = rr - sel * RQ;
            const int gx = gb * 4 + gxl;
            const float *rl = s_sub + (size_t)((2 * sel) * RQ + ry) * PS + kRX * gx;
            const float *rh = s_sub + (size_t)((2 * sel + 1) * RQ + ry) * PS + kRX * gx;
            float o[2 * kRX];
            if (F_ > 0) {
                constexpr int HW = F_ > 0 ? F_ / 2 : 1;
                constexpr int NW = (kRX + HW - 1 + 3) / 4;
                float wl[NW * 4], wh[NW * 4];
#pragma unroll
                for (int t = 0; t < NW; ++t) {
                    const float4 v = *reinterpret_cast<const float4 *>(rl + 4 * t);
                    wl[4 * t] = v.x; wl[4 * t + 1] = v.y; wl[4 * t + 2] = v.z; wl[4 * t + 3] = v.w;
                    const float4 u = *reinterpret_cast<const float4 *>(rh + 4 * t);
                    wh[4 * t] = u.x; wh[4 * t + 1] = u.y; wh[4 * t + 2] = u.z; wh[4 * t + 3] = u.w;
                }
#pragma unroll
                for (int c = 0; c < kRX; ++c) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float sl = 0.f, sh = 0.f;
#pragma unroll
                        for (int j = 0; j < HW; ++j) sl = mac<EXACT>(sl, taps.rec_lo[2 * j + e], wl[c + HW - 1 - j]);
#pragma unroll
                        for (int j = 0; j < HW; ++j) sh = mac<EXACT>(sh, taps.rec_hi[2 * j + e], wh[c + HW - 1 - j]);
                        o[2 * c + e] = __fadd_rn(sl, sh);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < kRX; ++c) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float sl = 0.f, sh = 0.f;
                        for (int j = 0; j < H; ++j) sl = mac<EXACT>(sl, taps.rec_lo[2 * j + e], rl[c + H - 1 - j]);
                        for (int j = 0; j < H; ++j) sh = mac<EXACT>(sh, taps.rec_hi[2 * j + e], rh[c + H - 1 - j]);
                        o[2 * c + e] = __fadd_rn(sl, sh);
                    }
                }
            }
            float *dst = s_mid + (size_t)(sel * RQ + ry) * PM + 2 * kRX * gx;
            *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    __syncthreads();

    // ---- axis -2 synthesis: (a, d) -> out rows 2q, 2q+1
    {
        constexpr int NGQ = kTQ / kRY;
        constexpr int OXW = 2 * kTP;
        float *dst = out.ptr + (size_t)blockIdx.z * out.plane_stride;
        for (int it = tid; it < NGQ * OXW; it += kNT) {
            const int gq = it / OXW, x = it - gq * OXW;
            const float *ca = s_mid + (size_t)(kRY * gq) * PM + x;
            const float *cd = s_mid + (size_t)(RQ + kRY * gq) * PM + x;
            float o[2 * kRY];
            if (F_ > 0) {
                constexpr int HW = F_ > 0 ? F_ / 2 : 1;
                float wa[kRY + HW - 1], wd[kRY + HW - 1];
#pragma unroll
                for (int t = 0; t < kRY + HW - 1; ++t) { wa[t] = ca[t * PM]; wd[t] = cd[t * PM]; }
#pragma unroll
                for (int c = 0; c < kRY; ++c) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float sl = 0.f, sh = 0.f;
#pragma unroll
                        for (int j = 0; j < HW; ++j) sl = mac<EXACT>(sl, taps.rec_lo[2 * j + e], wa[c + HW - 1 - j]);
#pragma unroll
                        for (int j = 0; j < HW; ++j) sh = mac<EXACT>(sh, taps.rec_hi[2 * j + e], wd[c + HW - 1 - j]);
                        o[2 * c + e] = __fadd_rn(sl, sh);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < kRY; ++c) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float sl = 0.f, sh = 0.f;
                        for (int j = 0; j < H; ++j) sl = mac<EXACT>(sl, taps.rec_lo[2 * j + e], ca[(c + H - 1 - j) * PM]);
                        for (int j = 0; j < H; ++j) sh = mac<EXACT>(sh, taps.rec_hi[2 * j + e], cd[(c + H - 1 - j) * PM]);
                        o[2 * c + e] = __fadd_rn(sl, sh);
                    }
                }
            }
            const int ox = 2 * p0 + x;
            if (ox < out.cols) {
                const int oyb = 2 * (q0 + kRY * gq);
#pragma unroll
                for (int k = 0; k < 2 * kRY; ++k)
                    if (oyb + k < out.rows) dst[(size_t)(oyb + k) * out.pitch + ox] = o[k];
            }
        }
    }
}

template <typename K>
void set_smem(K kernel, size_t bytes)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define B2S_FOR_STATIC_F(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(24) X(30)

template <int F, bool EXACT>
void launch_fwd_t(const B2sTaps &t, const B2sImg &in, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV,
                  const B2sImg &cD, int n_planes, cudaStream_t s)
{
    const FwdGeom g(t.F);
    const size_t bytes = g.smem_floats() * sizeof(float);
    set_smem(k_dwt_fwd<F, EXACT>, bytes);
    dim3 grid((cA.cols + kTX - 1) / kTX, (cA.rows + kTY - 1) / kTY, n_planes);
    k_dwt_fwd<F, EXACT><<<grid, kNT, bytes, s>>>(t, in, cA, cH, cV, cD);
}

template <int F, bool EXACT>
void launch_inv_t(const B2sTaps &t, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV, const B2sImg &cD,
                  const B2sImg &out, int n_planes, cudaStream_t s)
{
    const InvGeom g(t.F);
    const size_t bytes = g.smem_floats() * sizeof(float);
    set_smem(k_dwt_inv<F, EXACT>, bytes);
    dim3 grid((out.cols + 2 * kTP - 1) / (2 * kTP), (out.rows + 2 * kTQ - 1) / (2 * kTQ), n_planes);
    k_dwt_inv<F, EXACT><<<grid, kNT, bytes, s>>>(t, cA, cH, cV, cD, out);
}

}  // namespace

int b2s_dwt_max_smem(int F)
{
    const size_t a = FwdGeom(F).smem_floats() * sizeof(float), b = InvGeom(F).smem_floats() * sizeof(float);
    return (int)(a > b ? a : b);
}

void b2s_launch_dwt_fwd(const B2sTaps &t, const B2sImg &in, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV,
                        const B2sImg &cD, int n_planes, int exact, cudaStream_t s)
{
#define X(FF)                                                                                        \
    if (t.F == FF) {                                                                                 \
        if (exact) launch_fwd_t<FF, true>(t, in, cA, cH, cV, cD, n_planes, s);                       \
        else launch_fwd_t<FF, false>(t, in, cA, cH, cV, cD, n_planes, s);                            \
        return;                                                                                      \
    }
    B2S_FOR_STATIC_F(X)
#undef X
    if (exact) launch_fwd_t<0, true>(t, in, cA, cH, cV, cD, n_planes, s);
    else launch_fwd_t<0, false>(t, in, cA, cH, cV, cD, n_planes, s);
}

void b2s_launch_dwt_inv(const B2sTaps &t, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV, const B2sImg &cD,
                        const B2sImg &out, int n_planes, int exact, cudaStream_t s)
{
#define X(FF)                                                                                        \
    if (t.F == FF) {                                                                                 \
        if (exact) launch_inv_t<FF, true>(t, cA, cH, cV, cD, out, n_planes, s);                      \
        else launch_inv_t<FF, false>(t, cA, cH, cV, cD, out, n_planes, s);                           \
        return;                                                                                      \
    }
    B2S_FOR_STATIC_F(X)
#undef X
    if (exact) launch_inv_t<0, true>(t, cA, cH, cV, cD, out, n_planes, s);
    else launch_inv_t<0, false>(t, cA, cH, cV, cD, out, n_planes, s);
}
