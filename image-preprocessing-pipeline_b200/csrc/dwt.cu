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
// sm_100a design (the stage is FP32-issue bound, not HBM bound: 2F MACs per sample and pass)
//   * one CTA computes a TY x TX tile of all four sub-bands: the (2TY+F-2) x (2TX+F-2) input window is staged in
//     shared memory (128-bit coalesced loads; the half-sample symmetric extension is resolved at load time), the
//     axis -2 pass writes its two half-height images to a second shared buffer, the axis -1 pass reads them with
//     conflict-free 128-bit shared loads and stores 128-bit rows.  No intermediate goes to HBM.
//   * every thread owns R consecutive outputs and keeps the 2R+J-2 samples they need in registers, so one shared
//     load feeds ~6 MACs.
//   * the two outputs that share an input sample — (low, high) in the analysis, (even, odd) in the synthesis — are
//     computed together with packed f32x2 instructions (FFMA2 / FADD2 with a scalar-broadcast operand).  Packed
//     instructions have the same MAC rate as scalar ones on sm_100 but need half the issue slots, which leaves the
//     other half for the shared-memory loads, address arithmetic and stores: the FP32 pipe stays the only limiter.
//   * exact mode: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under -fmad=false, so the unfused product is
//     formed as fma(t, v, -0.0) with the -0.0 passed as a kernel argument (opaque to the compiler):
//     round(t*v - 0) == round(t*v) for every input, signed zeros included.
//   * filters longer than 20 taps (coif15 has 90) run the same code in chunks of J taps (taps are zero-padded to a
//     multiple of J; adding an exact zero product never changes a float32 sum), so registers stay bounded.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "b2s_internal.h"

namespace {

// ---- TMA (cp.async.bulk.tensor) tile loads ------------------------------------------------------------------------
// One elected thread fetches the whole input window of a tile with a single instruction: the copy engine walks the rows,
// fills everything outside the image with zeros and signals an mbarrier, so the other 255 threads spend no issue slots on
// address arithmetic (the LDGSTS loop was 13 % of the forward kernel's instructions, profiles/r02_*).  The inverse transform
// wants exactly those zeros; the forward transform's half-sample symmetric extension is patched in shared memory for the
// tiles that touch the border.  B2S_DWT_TMA=0 selects the LDGSTS loaders (kept for windows wider than a TMA box, 256).
typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmaEncodeFn tma_encode_fn()
{
    static TmaEncodeFn fn = [] {
        const char *e = getenv("B2S_DWT_TMA");
        if (e && atoi(e) == 0) return (TmaEncodeFn) nullptr;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return (TmaEncodeFn) nullptr; }
        return (TmaEncodeFn)p;
    }();
    return fn;
}
// 3-D map (column, row, plane) over the logical image: reads beyond rows / cols (also into the row padding) give zeros
bool make_tmap(CUtensorMap *m, const B2sImg &im, int n_planes, int box_w, int box_h)
{
    memset(m, 0, sizeof *m);
    TmaEncodeFn fn = tma_encode_fn();
    if (!fn || box_w > 256 || box_h > 256 || (reinterpret_cast<uintptr_t>(im.ptr) & 15) || (im.pitch & 3) || (im.plane_stride & 3))
        return false;
    cuuint64_t dims[3] = {(cuuint64_t)im.cols, (cuuint64_t)im.rows, (cuuint64_t)(n_planes > 0 ? n_planes : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)im.pitch * 4, (cuuint64_t)im.plane_stride * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, im.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(float *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int kMaxFp = 160;           // padded analysis filter length bound (B2S_MAX_TAPS rounded up to a chunk)
constexpr int kSmemPerSm = 227 * 1024;

// tile shapes.  Forward: TY x TX outputs per sub-band; inverse: TY x TX coefficient positions -> 2TY x 2TX outputs.
// R1 / R2 = consecutive outputs a thread computes from one register window in the first / second pass.
template <int TY_, int TX_, int NT_, int R1_, int R2_ = R1_>
struct Tile {
    static constexpr int TY = TY_, TX = TX_, NT = NT_, NW = NT_ / 32, R1 = R1_, R2 = R2_;
    static_assert(TX_ % 32 == 0 && NT_ % 32 == 0 && (R1_ == 4 || R1_ == 8) && (R2_ == 4 || R2_ == 8), "tile shape");
};

__host__ __device__ inline int round_up4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline int pitch_quads_odd(int n)  // >= n, multiple of 4, (p/4) odd
{
    int p = round_up4(n);
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}

__host__ __device__ inline int pitch_quads_4mod8(int n)  // >= n, multiple of 4, (p/4) % 8 == 4
{
    int p = round_up4(n);
    while (((p >> 2) & 7) != 4) p += 4;
    return p;
}

struct FwdTaps { float2 t[kMaxFp]; };                          // (dec_lo[j], dec_hi[j]), zero beyond F
struct InvTaps { float2 lo[kMaxFp / 2], hi[kMaxFp / 2]; };     // (rec[2j], rec[2j+1]), zero beyond F/2

// geometry of the shared-memory tiles for a padded filter length Fp (forward) / Hp = padded F/2 (inverse)
// shared-memory geometry: one input stage (the next tile is fetched into it while the second pass of the current tile
// runs out of the intermediate buffer) + the intermediate buffer
struct FwdGeom {
    int rin_y, rin_x, pin, pmid, stage_floats, mid_floats;
    __host__ __device__ FwdGeom(int Fp, int TY, int TX)
    {
        rin_y = 2 * TY + Fp - 2;
        rin_x = 2 * TX + Fp - 2;
        pin = round_up4(rin_x + 2) + 4;        // + alignment offset (0 or 2) of the tile's first column
        pmid = pitch_quads_odd(rin_x + 2);
        stage_floats = rin_y * pin;
        mid_floats = 2 * TY * pmid;
    }
    __host__ __device__ size_t smem_bytes() const { return sizeof(float) * ((size_t)stage_floats + mid_floats); }
};
struct InvGeom {
    int rq, rp, ps, pm, sub_floats, stage_floats, mid_floats;
    __host__ __device__ InvGeom(int Hp, int TQ, int TP, int R)
    {
        rq = TQ + Hp - 1;
        rp = TP + Hp - 1;
        // a quarter warp of the axis -1 pass reads 4 column groups (R floats apart) x 2 rows with 128-bit loads
        ps = R == 4 ? pitch_quads_4mod8(rp) : pitch_quads_odd(rp);
        pm = pitch_quads_odd(2 * TP);
        sub_floats = (rq * ps + 31) & ~31;   // every window starts on a 128-byte boundary (TMA destination)
        stage_floats = 4 * sub_floats;
        mid_floats = (2 * rq + 8) * pm;   // + slack rows: the last row group may look past TQ when TQ % R != 0
    }
    __host__ __device__ size_t smem_bytes() const { return sizeof(float) * ((size_t)stage_floats + mid_floats); }
};

__device__ __forceinline__ int sym_ext(int i, int n)  // half-sample symmetric extension, any i
{
    if (i >= 0 && i < n) return i;
    const int p = 2 * n;
    int t = i % p;
    if (t < 0) t += p;
    return t < n ? t : p - 1 - t;
}

// asynchronous global -> shared copies (LDGSTS): a thread issues its whole share of the tile without waiting
__device__ __forceinline__ void cp_async16(float *dst, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// (t.x, t.y) * v, rounded once per lane: see the header for why this is not __fmul2_rn
__device__ __forceinline__ float2 mul2_exact(float2 t, float v, float2 nz) { return __ffma2_rn(t, make_float2(v, v), nz); }

__device__ __forceinline__ float mac1_exact(float acc, float a, float b) { return __fadd_rn(acc, __fmul_rn(a, b)); }

// "acc (+)= t * (v, v)" on an output pair.  Exact: unfused multiply then add (the reference's rounding); fast: one FMA.
// (Measured on B200: scalar FMUL/FADD or FFMA variants of the same loops are no faster than the packed ones.)
enum { kExact = 0, kFast = 1 };
template <int MODE>
__device__ __forceinline__ void mac_pair(float2 &acc, float2 t, float v, float2 nz, bool first)
{
    if (MODE == kExact) {
        const float2 p = mul2_exact(t, v, nz);
        acc = first ? p : __fadd2_rn(acc, p);
    } else {
        acc = first ? __fmul2_rn(t, make_float2(v, v)) : __ffma2_rn(t, make_float2(v, v), acc);
    }
}

// ---- R consecutive analysis outputs from a register window -----------------------------------------------------
// out[r] = sum_j taps[j] * x[2r + Fp-1 - j], j ascending; x = `base` with element stride `stride` (VEC: stride 1,
// 16-byte aligned base, 128-bit loads).
template <int R, int J, bool MULTI, int MODE, bool VEC>
__device__ __forceinline__ void analysis_run(const float *base, int stride, const float2 *__restrict__ taps, int nch,
                                             int Fp, float2 nz, float2 (&acc)[R])
{
    constexpr int W = 2 * R + J - 2;
    constexpr int W4 = (W + 3) / 4;
    float w[W4 * 4];
    if (!MULTI) {
        if (VEC) {
#pragma unroll
            for (int k = 0; k < W4; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(base + 4 * k);
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; ++k) w[k] = base[k * stride];
        }
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            const float2 t = taps[jj];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                mac_pair<MODE>(acc[r], t, w[2 * r + J - 1 - jj], nz, jj == 0);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
        for (int c = 0; c < nch; ++c) {
            const float *b = base + (Fp - (c + 1) * J) * stride;   // multiple of 4 elements (J % 4 == 0)
            if (VEC) {
#pragma unroll
                for (int k = 0; k < W4; ++k) {
                    const float4 v = *reinterpret_cast<const float4 *>(b + 4 * k);
                    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < W; ++k) w[k] = b[k * stride];
            }
            const float2 *tc = taps + c * J;
#pragma unroll
            for (int jj = 0; jj < J; ++jj) {
                const float2 t = tc[jj];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    mac_pair<MODE>(acc[r], t, w[2 * r + J - 1 - jj], nz, false);
                }
            }
        }
    }
}

// analysis output whose window overhangs the right edge (i = 2o+1 >= n): pywt visits the reflected part with the
// filter index descending from i-n to 0, then the in-range part ascending.  `w0` points at the extended sample that
// pairs with tap 0; higher taps pair with samples `step` floats lower.
__device__ __forceinline__ float2 analysis_right_edge(const FwdTaps &taps, int F, const float *w0, int step, int over)
{
    float lo = 0.f, hi = 0.f;
#pragma unroll 1
    for (int t = 0; t < F; ++t) {
        const int j = (t <= over) ? (over - t) : t;
        const float2 f = taps.t[j];
        const float v = w0[-j * step];
        lo = mac1_exact(lo, f.x, v);
        hi = mac1_exact(hi, f.y, v);
    }
    return make_float2(lo, hi);
}

// ---- R consecutive synthesis positions -> R (even, odd) output pairs ---------------------------------------------
// pair[cc] = sum_j lo[j] * a[cc + Hp-1 - j]  (+)  sum_j hi[j] * d[cc + Hp-1 - j], each sum j ascending, one final add.
template <int R, int JH, bool MULTI, int MODE, bool VEC>
__device__ __forceinline__ void synthesis_run(const float *pa, const float *pd, int stride, const float2 *__restrict__ tlo,
                                              const float2 *__restrict__ thi, int nch, int Hp, float2 nz, float2 (&out)[R])
{
    constexpr int W = R + JH - 1;
    constexpr int W4 = (W + 3) / 4;
    float wa[W4 * 4], wd[W4 * 4];
    float2 sl[R], sh[R];
    if (!MULTI) {
        if (VEC) {
#pragma unroll
            for (int k = 0; k < W4; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(pa + 4 * k);
                wa[4 * k] = v.x; wa[4 * k + 1] = v.y; wa[4 * k + 2] = v.z; wa[4 * k + 3] = v.w;
                const float4 u = *reinterpret_cast<const float4 *>(pd + 4 * k);
                wd[4 * k] = u.x; wd[4 * k + 1] = u.y; wd[4 * k + 2] = u.z; wd[4 * k + 3] = u.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; ++k) { wa[k] = pa[k * stride]; wd[k] = pd[k * stride]; }
        }
#pragma unroll
        for (int jj = 0; jj < JH; ++jj) {
            const float2 fl = tlo[jj], fh = thi[jj];
#pragma unroll
            for (int cc = 0; cc < R; ++cc) {
                mac_pair<MODE>(sl[cc], fl, wa[cc + JH - 1 - jj], nz, jj == 0);
                mac_pair<MODE>(sh[cc], fh, wd[cc + JH - 1 - jj], nz, jj == 0);
            }
        }
    } else {
#pragma unroll
        for (int cc = 0; cc < R; ++cc) { sl[cc] = make_float2(0.f, 0.f); sh[cc] = make_float2(0.f, 0.f); }
        for (int c = 0; c < nch; ++c) {
            const int b = (Hp - (c + 1) * JH) * stride;
            if (VEC) {
#pragma unroll
                for (int k = 0; k < W4; ++k) {
                    const float4 v = *reinterpret_cast<const float4 *>(pa + b + 4 * k);
                    wa[4 * k] = v.x; wa[4 * k + 1] = v.y; wa[4 * k + 2] = v.z; wa[4 * k + 3] = v.w;
                    const float4 u = *reinterpret_cast<const float4 *>(pd + b + 4 * k);
                    wd[4 * k] = u.x; wd[4 * k + 1] = u.y; wd[4 * k + 2] = u.z; wd[4 * k + 3] = u.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < W; ++k) { wa[k] = pa[b + k * stride]; wd[k] = pd[b + k * stride]; }
            }
            const float2 *cl = tlo + c * JH, *ch = thi + c * JH;
#pragma unroll
            for (int jj = 0; jj < JH; ++jj) {
                const float2 fl = cl[jj], fh = ch[jj];
#pragma unroll
                for (int cc = 0; cc < R; ++cc) {
                    mac_pair<MODE>(sl[cc], fl, wa[cc + JH - 1 - jj], nz, false);
                    mac_pair<MODE>(sh[cc], fh, wd[cc + JH - 1 - jj], nz, false);
                }
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < R; ++cc) out[cc] = __fadd2_rn(sl[cc], sh[cc]);
}

// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
    B2sImg in, cA, cH, cV, cD;
    int F, Fp, nch;
    int tiles_x, tiles_y, n_tiles;
    int grid3d;      // launched as (tiles_x, tiles_y, planes): one tile per CTA
    int use_tma;     // the input window arrives through the tensor map
    float negzero;   // -0.0f, deliberately a run-time value (see mul2_exact)
};

struct TileCoord { int plane, ty, tx; };
__device__ __forceinline__ TileCoord tile_of(int t, int tiles_x, int tiles_xy)
{
    TileCoord c;
    c.plane = t / tiles_xy;
    const int r2 = t - c.plane * tiles_xy;
    c.ty = r2 / tiles_x;
    c.tx = r2 - c.ty * tiles_x;
    return c;
}

// Persistent CTAs: each walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The input window of the next tile is
// copied (cp.async) into the input stage while the axis -1 pass of the current tile runs out of the intermediate buffer.
template <class T, int J, bool MULTI, int MODE>
__global__ void __launch_bounds__(T::NT) k_dwt_fwd(const __grid_constant__ FwdTaps taps, const FwdArgs a,
                                                   const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    // TMA destinations must sit on a 128-byte boundary of the shared window: align explicitly (the launch adds 128 bytes)
    float *smem = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    constexpr int TY = T::TY, TX = T::TX, NT = T::NT, NW = T::NW, RY = T::R1, RX = T::R2;
    constexpr bool EXACT = MODE == kExact;
    const int Fp = MULTI ? a.Fp : J;
    const FwdGeom g(Fp, TY, TX);
    const int PIN = g.pin, PM = g.pmid;
    float *s_in = smem;
    float *s_mid = smem + g.stage_floats;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ny = a.in.rows, nx = a.in.cols;
    const int my = a.cA.rows, mx = a.cA.cols;
    const int off = (2 - Fp) & 3;             // tile column origin 2*ox0 + 2 - Fp modulo 4 (ox0 is a multiple of 32)
    const int c4n = (g.rin_x + off + 3) >> 2;
    const float2 nz = make_float2(a.negzero, a.negzero);
    const int tiles_xy = a.tiles_x * a.tiles_y;

    // stage the input window of tile t (half-sample symmetric extension resolved here); asynchronous
    // 3-D launch (one tile per CTA): coordinates come from blockIdx, no integer division on the critical path
    const bool grid3d = gridDim.y > 1 || gridDim.z > 1 || a.grid3d;
    auto coord_of = [&](int t) -> TileCoord {
        if (grid3d) { TileCoord c; c.plane = blockIdx.z; c.ty = blockIdx.y; c.tx = blockIdx.x; return c; }
        return tile_of(t, a.tiles_x, tiles_xy);
    };
    auto issue_load = [&](int t) {
        const TileCoord c = coord_of(t);
        const int gy0 = 2 * c.ty * TY + 2 - Fp;            // image row of shared row 0
        const int gx0a = (2 * c.tx * TX + 2 - Fp) & ~3;    // 16-byte aligned image column of shared column 0
        const float *src = a.in.ptr + (size_t)c.plane * a.in.plane_stride;
        if (gy0 >= 0 && gy0 + g.rin_y <= ny && gx0a >= 0 && gx0a + 4 * c4n <= nx) {
            // interior tile: every quad is one aligned 16-byte copy; (row, quad) advance incrementally
            const int nq = g.rin_y * c4n;
            int r = tid / c4n, c4 = tid - r * c4n;
            const int dr = NT / c4n, dc = NT - dr * c4n;
            const float *sp = src + (size_t)(gy0 + r) * a.in.pitch + gx0a + 4 * c4;
            float *dp = s_in + r * PIN + 4 * c4;
            const int s_step = dr * a.in.pitch + 4 * dc, d_step = dr * PIN + 4 * dc;
            const int s_wrap = a.in.pitch - 4 * c4n, d_wrap = PIN - 4 * c4n;
            for (int q = tid; q < nq; q += NT) {
                cp_async16(dp, sp);
                sp += s_step; dp += d_step; c4 += dc;
                if (c4 >= c4n) { c4 -= c4n; sp += s_wrap; dp += d_wrap; }
            }
        } else {
            // edge tile: only samples within Fp of the image can reach a stored output; the rest of the window is left
            // as it is (it only feeds outputs beyond the sub-band, which are never stored)
            for (int r = warp; r < g.rin_y; r += NW) {
                const int gy = gy0 + r;
                if (gy < -Fp || gy >= ny + Fp) continue;
                const float *srow = src + (size_t)sym_ext(gy, ny) * a.in.pitch;
                float *drow = s_in + r * PIN;
                for (int c4 = lane; c4 < c4n; c4 += 32) {
                    const int gx = gx0a + 4 * c4;
                    if (gx >= 0 && gx + 3 < nx) {
                        cp_async16(drow + 4 * c4, srow + gx);
                    } else if (gx + 3 >= -Fp && gx < nx + Fp) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) cp_async4(drow + 4 * c4 + k, srow + sym_ext(gx + k, nx));
                    }
                }
            }
        }
    };

    int t = grid3d ? 0 : blockIdx.x;
    const int t_end = grid3d ? 1 : a.n_tiles, t_step = grid3d ? 1 : gridDim.x;
    const bool tma = a.use_tma != 0;
    unsigned phase = 0;
    // TMA: one thread arms the barrier with the byte count of the box and starts the copy; everybody waits on the barrier
    auto tma_issue = [&](int tt) {
        if (tid == 0) {
            const TileCoord c = coord_of(tt);
            mbar_expect(&s_bar, (unsigned)(g.rin_y * PIN * sizeof(float)));
            tma_load_3d(s_in, &tmap, &s_bar, (2 * c.tx * TX + 2 - Fp) & ~3, 2 * c.ty * TY + 2 - Fp, c.plane);
        }
    };
    // border tile: the copy engine wrote zeros outside the image; replace those within reach of a stored output by the
    // half-sample symmetric extension (read from global memory: any image size, any number of reflections)
    auto tma_fixup = [&](int tt) {
        const TileCoord c = coord_of(tt);
        const int gy0 = 2 * c.ty * TY + 2 - Fp, gx0a = (2 * c.tx * TX + 2 - Fp) & ~3;
        if (!(gy0 < 0 || gy0 + g.rin_y > ny || gx0a < 0 || gx0a + 4 * c4n > nx)) return;
        const float *src = a.in.ptr + (size_t)c.plane * a.in.plane_stride;
        const int wq = 4 * c4n;
        for (int r = warp; r < g.rin_y; r += NW) {
            const int gy = gy0 + r;
            if (gy < -Fp || gy >= ny + Fp) continue;
            const bool row_in = gy >= 0 && gy < ny;
            const float *srow = src + (size_t)sym_ext(gy, ny) * a.in.pitch;
            float *drow = s_in + r * PIN;
            if (row_in) {   // only the columns left / right of the image
                for (int cc = lane; cc < 2 * Fp + 8; cc += 32) {
                    const int gx = cc < Fp + 4 ? -1 - cc : nx + (cc - Fp - 4);
                    const int sc = gx - gx0a;
                    if (sc >= 0 && sc < wq && gx >= -Fp && gx < nx + Fp) drow[sc] = srow[sym_ext(gx, nx)];
                }
            } else {
                for (int sc = lane; sc < wq; sc += 32) {
                    const int gx = gx0a + sc;
                    if (gx >= -Fp && gx < nx + Fp) drow[sc] = srow[sym_ext(gx, nx)];
                }
            }
        }
    };
    if (tma) {
        if (tid == 0) mbar_init(&s_bar);
        __syncthreads();
        if (t < t_end) tma_issue(t);
    } else if (t < t_end) issue_load(t);
    while (t < t_end) {
        if (tma) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            tma_fixup(t);
        } else cp_async_wait_all();
        __syncthreads();   // tile t has landed; every warp is done with s_mid of the previous tile

        const TileCoord c = coord_of(t);
        const int oy0 = c.ty * TY, ox0 = c.tx * TX;
        // does this tile hold outputs whose window overhangs the bottom / right edge (the reference visits their taps
        // in another order, see analysis_right_edge)?
        const bool edge_y = 2 * (oy0 + TY) - 1 >= ny, edge_x = 2 * (ox0 + TX) - 1 >= nx;

        // ---- axis -2 (rows index) pass: s_in -> s_mid[0..TY) = low-pass rows, s_mid[TY..2TY) = high-pass rows
        {
            const int ncg = (g.rin_x + 31) >> 5;
            constexpr int NGY = (TY + RY - 1) / RY;
            for (int wi = warp; wi < NGY * ncg; wi += NW) {
                const int gy = wi / ncg;
                const int col_idx = (wi - gy * ncg) * 32 + lane;
                if (col_idx >= g.rin_x) continue;
                const float *col = s_in + (2 * RY * gy) * PIN + col_idx + off;
                float2 acc[RY];
                analysis_run<RY, J, MULTI, MODE, false>(col, PIN, taps.t, a.nch, Fp, nz, acc);
                if (EXACT && edge_y) {
#pragma unroll
                    for (int r = 0; r < RY; ++r) {
                        const int i = 2 * (oy0 + RY * gy + r) + 1;
                        if (i >= ny && i - ny <= a.F - 2)
                            acc[r] = analysis_right_edge(taps, a.F, col + (2 * r + Fp - 1) * PIN, PIN, i - ny);
                    }
                }
#pragma unroll
                for (int r = 0; r < RY; ++r) {
                    if (TY % RY == 0 || RY * gy + r < TY) {
                        s_mid[(RY * gy + r) * PM + col_idx] = acc[r].x;
                        s_mid[(TY + RY * gy + r) * PM + col_idx] = acc[r].y;
                    }
                }
            }
        }
        __syncthreads();   // s_mid complete; s_in is free again

        const int tn = t + t_step;
        if (tn < t_end) { if (tma) tma_issue(tn); else issue_load(tn); }

        // ---- axis -1 pass.  RX = 4: a warp covers 8 rows x 4 column groups; RX = 8: 16 rows x 2 column groups.  Either way
        //      the 128-bit window loads of a quarter warp hit 8 different bank quads (PM/4 odd).
        {
            constexpr int GPW = RX == 4 ? 4 : 2;          // column groups per warp
            constexpr int RPW = 32 / GPW;                // rows per warp
            const int gxl = lane % GPW, rsub = lane / GPW;
            constexpr int N_GB = TX / RX / GPW;
            constexpr int N_RB = (2 * TY + RPW - 1) / RPW;
            float *pA = a.cA.ptr + (size_t)c.plane * a.cA.plane_stride;
            float *pH = a.cH.ptr + (size_t)c.plane * a.cH.plane_stride;
            float *pV = a.cV.ptr + (size_t)c.plane * a.cV.plane_stride;
            float *pD = a.cD.ptr + (size_t)c.plane * a.cD.plane_stride;
            for (int wi = warp; wi < N_RB * N_GB; wi += NW) {
                const int rb = wi / N_GB, gb = wi - rb * N_GB;
                const int rm = rb * RPW + rsub;
                if ((2 * TY) % RPW != 0 && rm >= 2 * TY) continue;
                const int gx = gb * GPW + gxl;
                const float *row = s_mid + rm * PM + 2 * RX * gx;
                float2 acc[RX];
                analysis_run<RX, J, MULTI, MODE, true>(row, 1, taps.t, a.nch, Fp, nz, acc);
                const int oxb = ox0 + RX * gx;
                if (EXACT && edge_x) {
#pragma unroll
                    for (int cc = 0; cc < RX; ++cc) {
                        const int i = 2 * (oxb + cc) + 1;
                        if (i >= nx && i - nx <= a.F - 2)
                            acc[cc] = analysis_right_edge(taps, a.F, row + 2 * cc + Fp - 1, 1, i - nx);
                    }
                }
                const bool low_rows = rm < TY;
                const int oy = oy0 + (low_rows ? rm : rm - TY);
                if (oy < my) {   // pitch is a multiple of 4, so a whole float4 always fits when its first column does
                    const size_t o = (size_t)oy * a.cA.pitch + oxb;     // the four sub-bands share pitch
                    float *d0 = (low_rows ? pA : pH) + o, *d1 = (low_rows ? pV : pD) + o;
#pragma unroll
                    for (int h = 0; h < RX / 4; ++h) {
                        if (oxb + 4 * h < mx) {
                            *reinterpret_cast<float4 *>(d0 + 4 * h) =
                                make_float4(acc[4 * h].x, acc[4 * h + 1].x, acc[4 * h + 2].x, acc[4 * h + 3].x);
                            *reinterpret_cast<float4 *>(d1 + 4 * h) =
                                make_float4(acc[4 * h].y, acc[4 * h + 1].y, acc[4 * h + 2].y, acc[4 * h + 3].y);
                        }
                    }
                }
            }
        }
        t = tn;
    }
}

// ------------------------------------------------------------------------------------------------ inverse
struct InvArgs {
    B2sImg cA, cH, cV, cD, out;
    int H, Hp, nch;
    int tiles_x, tiles_y, n_tiles;
    int grid3d;
    int use_tma;
    float negzero;
};
struct InvMaps { CUtensorMap m[4]; };   // cA, cV, cH, cD (the order of the shared-memory windows)

template <class T, int JH, bool MULTI, int MODE>
__global__ void __launch_bounds__(T::NT) k_dwt_inv(const __grid_constant__ InvTaps taps, const InvArgs a,
                                                   const __grid_constant__ InvMaps maps)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    float *smem = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    constexpr int TQ = T::TY, TP = T::TX, NT = T::NT, NW = T::NW, RX = T::R1, RY = T::R2;
    const int Hp = MULTI ? a.Hp : JH;
    const InvGeom g(Hp, TQ, TP, RX);
    const int PS = g.ps, PM = g.pm, RQ = g.rq;
    const int sub_floats = g.sub_floats;
    float *s_sub = smem;                       // [cA, cV, cH, cD][RQ][PS]
    float *s_mid = smem + g.stage_floats;      // [2][RQ][PM]: a (from cA,cV), d (from cH,cD)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my = a.cH.rows, mx = a.cH.cols;
    const float2 nz = make_float2(a.negzero, a.negzero);
    const int tiles_xy = a.tiles_x * a.tiles_y;
    const int c4n = (g.rp + 3) >> 2;

    // stage the four coefficient windows of tile t; everything outside the sub-band is zero (it only meets zero taps
    // or outputs that are not stored)
    const bool grid3d = gridDim.y > 1 || gridDim.z > 1 || a.grid3d;
    auto coord_of = [&](int t) -> TileCoord {
        if (grid3d) { TileCoord c; c.plane = blockIdx.z; c.ty = blockIdx.y; c.tx = blockIdx.x; return c; }
        return tile_of(t, a.tiles_x, tiles_xy);
    };
    auto issue_load = [&](int t) {
        const TileCoord c = coord_of(t);
        const int cy0 = c.ty * TQ + a.H - Hp, cx0 = c.tx * TP + a.H - Hp;   // coefficient coordinates of shared (0, 0)
        const bool aligned = (cx0 & 3) == 0;
        if (aligned && cy0 >= 0 && cy0 + RQ <= my && cx0 >= 0 && cx0 + 4 * c4n <= mx) {
            const int nq = RQ * c4n;
            const int r0 = tid / c4n, c40 = tid - r0 * c4n;
            const int dr = NT / c4n, dc = NT - dr * c4n;
#pragma unroll
            for (int sb = 0; sb < 4; ++sb) {
                const B2sImg &im = sb == 0 ? a.cA : (sb == 1 ? a.cV : (sb == 2 ? a.cH : a.cD));
                const float *sp = im.ptr + (size_t)c.plane * im.plane_stride + (size_t)(cy0 + r0) * im.pitch + cx0 + 4 * c40;
                float *dp = s_sub + sb * sub_floats + r0 * PS + 4 * c40;
                const int s_step = dr * im.pitch + 4 * dc, d_step = dr * PS + 4 * dc;
                const int s_wrap = im.pitch - 4 * c4n, d_wrap = PS - 4 * c4n;
                int c4 = c40;
                for (int q = tid; q < nq; q += NT) {
                    cp_async16(dp, sp);
                    sp += s_step; dp += d_step; c4 += dc;
                    if (c4 >= c4n) { c4 -= c4n; sp += s_wrap; dp += d_wrap; }
                }
            }
        } else {
            for (int r = warp; r < 4 * RQ; r += NW) {
                const int sb = r / RQ, ry = r - sb * RQ;
                const int y = cy0 + ry;
                float *drow = s_sub + sb * sub_floats + ry * PS;
                const bool row_ok = y >= 0 && y < my;
                const B2sImg &im = sb == 0 ? a.cA : (sb == 1 ? a.cV : (sb == 2 ? a.cH : a.cD));
                const float *srow = im.ptr + (size_t)c.plane * im.plane_stride + (size_t)(row_ok ? y : 0) * im.pitch;
                for (int c4 = lane; c4 < c4n; c4 += 32) {
                    const int x = cx0 + 4 * c4;
                    if (row_ok && aligned && x >= 0 && x + 3 < mx) {
                        cp_async16(drow + 4 * c4, srow + x);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (row_ok && x + k >= 0 && x + k < mx) cp_async4(drow + 4 * c4 + k, srow + x + k);
                            else drow[4 * c4 + k] = 0.f;
                        }
                    }
                }
            }
        }
    };

    int t = grid3d ? 0 : blockIdx.x;
    const int t_end = grid3d ? 1 : a.n_tiles, t_step = grid3d ? 1 : gridDim.x;
    const bool tma = a.use_tma != 0;
    unsigned phase = 0;
    // TMA: four boxes (cA, cV, cH, cD), one barrier; coefficients outside the sub-band arrive as the zeros the synthesis wants
    auto tma_issue = [&](int tt) {
        if (tid == 0) {
            const TileCoord c = coord_of(tt);
            const int cy0 = c.ty * TQ + a.H - Hp, cx0 = c.tx * TP + a.H - Hp;
            mbar_expect(&s_bar, (unsigned)(4 * RQ * PS * sizeof(float)));
#pragma unroll
            for (int sb = 0; sb < 4; ++sb) tma_load_3d(s_sub + sb * sub_floats, &maps.m[sb], &s_bar, cx0, cy0, c.plane);
        }
    };
    if (tma) {
        if (tid == 0) mbar_init(&s_bar);
        __syncthreads();
        if (t < t_end) tma_issue(t);
    } else if (t < t_end) issue_load(t);
    while (t < t_end) {
        if (tma) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
        } else cp_async_wait_all();
        __syncthreads();   // tile t has landed; every warp is done with s_mid of the previous tile

        const TileCoord c = coord_of(t);
        const int q0 = c.ty * TQ, p0 = c.tx * TP;

        // ---- axis -1 synthesis: (cA,cV) -> a, (cH,cD) -> d   [idwtn handles the last axis first]
        //      a warp covers 8 rows x 4 groups of RX coefficient positions (PS/4 odd: conflict-free 128-bit loads)
        {
            const int gxl = lane & 3, rsub = lane >> 2;
            constexpr int N_GB = TP / RX / 4;
            const int n_rb = (2 * RQ + 7) >> 3;
            for (int wi = warp; wi < n_rb * N_GB; wi += NW) {
                const int rb = wi / N_GB, gb = wi - rb * N_GB;
                const int rr = rb * 8 + rsub;
                if (rr >= 2 * RQ) continue;
                const int sel = rr >= RQ ? 1 : 0;
                const int ry = rr - sel * RQ;
                const int gx = gb * 4 + gxl;
                const float *rl = s_sub + (2 * sel) * sub_floats + ry * PS + RX * gx;
                const float *rh = rl + sub_floats;
                float2 o[RX];  // (even, odd) output pairs
                synthesis_run<RX, JH, MULTI, MODE, true>(rl, rh, 1, taps.lo, taps.hi, a.nch, Hp, nz, o);
                float *dst = s_mid + (size_t)(sel * RQ + ry) * PM + 2 * RX * gx;
#pragma unroll
                for (int h = 0; h < RX / 2; ++h)
                    *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(o[2 * h].x, o[2 * h].y, o[2 * h + 1].x, o[2 * h + 1].y);
            }
        }
        __syncthreads();   // s_mid complete; s_sub is free again

        const int tn = t + t_step;
        if (tn < t_end) { if (tma) tma_issue(tn); else issue_load(tn); }

        // ---- axis -2 synthesis: (a, d) -> out rows 2q, 2q+1
        {
            constexpr int NGQ = (TQ + RY - 1) / RY;
            constexpr int NCG = (2 * TP) / 32;
            float *dstp = a.out.ptr + (size_t)c.plane * a.out.plane_stride;
            for (int wi = warp; wi < NGQ * NCG; wi += NW) {
                const int gq = wi / NCG;
                const int x = (wi - gq * NCG) * 32 + lane;
                const float *ca = s_mid + (size_t)(RY * gq) * PM + x;
                const float *cd = ca + (size_t)RQ * PM;
                float2 o[RY];  // (row 2q, row 2q+1)
                synthesis_run<RY, JH, MULTI, MODE, false>(ca, cd, PM, taps.lo, taps.hi, a.nch, Hp, nz, o);
                const int ox = 2 * p0 + x;
                if (ox < a.out.cols) {
                    const int oyb = 2 * (q0 + RY * gq);
#pragma unroll
                    for (int k = 0; k < RY; ++k) {
                        if (TQ % RY != 0 && RY * gq + k >= TQ) continue;
                        if (oyb + 2 * k < a.out.rows) dstp[(size_t)(oyb + 2 * k) * a.out.pitch + ox] = o[k].x;
                        if (oyb + 2 * k + 1 < a.out.rows) dstp[(size_t)(oyb + 2 * k + 1) * a.out.pitch + ox] = o[k].y;
                    }
                }
            }
        }
        t = tn;
    }
}

// ================================================================================================ long filters
// F > 40 (coif7..17, db21..38, sym21+): the fused 2-D tile above would carry a (2TX+F-2)/(2TX) halo through its first
// pass and fit one CTA per SM.  Here each axis is its own kernel: no output is computed twice, three CTAs fit an SM, and
// the half-height / half-width intermediates (lo, hi / a, d) go through a per-slot scratch buffer (L2 / HBM), which
// is cheap next to the 2F multiply-adds per sample.  Arithmetic is the chunked exact path of analysis_run /
// synthesis_run.  The zero taps that pad the filter to a whole number of chunks sit in FRONT of the real taps (a sum
// that starts at +0 is unchanged by adding +-0 products first), chosen so that every tile window starts on a 16-byte
// boundary of the source rows.
struct FwdLongArgs {
    B2sImg in, lo, hi, cA, cH, cV, cD;
    int F, Fp, nch, shift;   // shift = number of leading zero taps
    float negzero;
};
struct InvLongArgs {
    B2sImg cA, cH, cV, cD, a, d, out;
    int H, Hp, nch, shift;
    float negzero;
};

__device__ __forceinline__ float2 analysis_right_edge_p(const float2 *__restrict__ taps, int F, const float *w0, int step, int over)
{
    float lo = 0.f, hi = 0.f;
#pragma unroll 1
    for (int t = 0; t < F; ++t) {
        const int j = (t <= over) ? (over - t) : t;
        const float2 f = taps[j];
        const float v = w0[-j * step];
        lo = mac1_exact(lo, f.x, v);
        hi = mac1_exact(hi, f.y, v);
    }
    return make_float2(lo, hi);
}

constexpr int kLongNT = 256;
// forward, axis -2: TY output rows x TC columns per CTA; a thread owns RY consecutive output rows of one column
constexpr int kFyTY = 64, kFyTC = 64, kFyR = 8;
// forward, axis -1: TR rows of lo and of hi x TX outputs per CTA
constexpr int kFxTR = 16, kFxTX = 128, kFxR = 8;
// inverse, axis -1: TR coefficient rows (both pairs) x TP coefficient columns -> 2 TP outputs
constexpr int kIxTR = 8, kIxTP = 128, kIxR = 4;
// inverse, axis -2: TQ coefficient rows x TC columns -> 2 TQ output rows
constexpr int kIyTQ = 64, kIyTC = 64, kIyR = 4;

template <int J, int MODE>
__global__ void __launch_bounds__(kLongNT) k_dwt_fwd_y_long(const __grid_constant__ FwdTaps taps, const FwdLongArgs a)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int TY = kFyTY, TC = kFyTC, RY = kFyR, NT = kLongNT, NW = NT / 32;
    const int Fp = a.Fp;
    const int rin_y = 2 * TY + Fp - 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ny = a.in.rows, nx = a.in.cols, my = a.lo.rows;
    const int plane = blockIdx.z, oy0 = blockIdx.y * TY, x0 = blockIdx.x * TC;
    const int gy0 = 2 * oy0 + 2 - Fp + a.shift;
    const float2 nz = make_float2(a.negzero, a.negzero);
    const float *src = a.in.ptr + (size_t)plane * a.in.plane_stride;
    for (int idx = tid; idx < rin_y * (TC / 4); idx += NT) {
        const int r = idx / (TC / 4), q = idx % (TC / 4);
        const int gx = x0 + 4 * q;
        if (gx < a.in.pitch) cp_async16(smem + r * TC + 4 * q, src + (size_t)sym_ext(gy0 + r, ny) * a.in.pitch + gx);
    }
    cp_async_wait_all();
    __syncthreads();
    const bool edge_y = 2 * (oy0 + TY) - 1 >= ny;
    float *plo = a.lo.ptr + (size_t)plane * a.lo.plane_stride, *phi = a.hi.ptr + (size_t)plane * a.hi.plane_stride;
    constexpr int NCG = TC / 32;
    for (int wi = warp; wi < (TY / RY) * NCG; wi += NW) {
        const int gy = wi / NCG, col_idx = (wi % NCG) * 32 + lane;
        const float *col = smem + (2 * RY * gy) * TC + col_idx;
        float2 acc[RY];
        analysis_run<RY, J, true, MODE, false>(col, TC, taps.t, a.nch, Fp, nz, acc);
        if (MODE == kExact && edge_y) {
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                const int i = 2 * (oy0 + RY * gy + r) + 1;
                if (i >= ny && i - ny <= a.F - 2)
                    acc[r] = analysis_right_edge_p(taps.t + a.shift, a.F, col + (2 * r + Fp - 1 - a.shift) * TC, TC, i - ny);
            }
        }
        const int x = x0 + col_idx;
        if (x < nx) {
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                const int oy = oy0 + RY * gy + r;
                if (oy < my) {
                    plo[(size_t)oy * a.lo.pitch + x] = acc[r].x;
                    phi[(size_t)oy * a.hi.pitch + x] = acc[r].y;
                }
            }
        }
    }
}

template <int J, int MODE>
__global__ void __launch_bounds__(kLongNT) k_dwt_fwd_x_long(const __grid_constant__ FwdTaps taps, const FwdLongArgs a)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int TR = kFxTR, TX = kFxTX, RX = kFxR, NT = kLongNT, NW = NT / 32;
    const int Fp = a.Fp;
    const int rin_x = 2 * TX + Fp - 2;
    const int PM = pitch_quads_odd(rin_x + 2);
    const int c4n = (rin_x + 3) >> 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nx = a.lo.cols, my = a.cA.rows, mx = a.cA.cols;
    const int plane = blockIdx.z, oy0 = blockIdx.y * TR, ox0 = blockIdx.x * TX;
    const int gx0 = 2 * ox0 + 2 - Fp + a.shift;
    const bool aligned = (gx0 & 3) == 0;
    const float2 nz = make_float2(a.negzero, a.negzero);
    const float *slo = a.lo.ptr + (size_t)plane * a.lo.plane_stride, *shi = a.hi.ptr + (size_t)plane * a.hi.plane_stride;
    for (int idx = tid; idx < 2 * TR * c4n; idx += NT) {
        const int rm = idx / c4n, c4 = idx - rm * c4n;
        const int oy = oy0 + (rm % TR);
        if (oy >= my) continue;   // rows beyond the sub-band only feed outputs that are not stored
        const float *srow = (rm < TR ? slo : shi) + (size_t)oy * a.lo.pitch;
        const int gx = gx0 + 4 * c4;
        float *dst = smem + rm * PM + 4 * c4;
        if (aligned && gx >= 0 && gx + 3 < nx) {
            cp_async16(dst, srow + gx);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async4(dst + k, srow + sym_ext(gx + k, nx));
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const bool edge_x = 2 * (ox0 + TX) - 1 >= nx;
    constexpr int GPW = 2, RPW = 16;           // RX = 8: a warp covers 16 rows x 2 column groups (PM/4 odd: conflict-free)
    const int gxl = lane % GPW, rsub = lane / GPW;
    constexpr int N_GB = TX / RX / GPW;
    constexpr int N_RB = (2 * TR) / RPW;
    float *pA = a.cA.ptr + (size_t)plane * a.cA.plane_stride;
    float *pH = a.cH.ptr + (size_t)plane * a.cH.plane_stride;
    float *pV = a.cV.ptr + (size_t)plane * a.cV.plane_stride;
    float *pD = a.cD.ptr + (size_t)plane * a.cD.plane_stride;
    for (int wi = warp; wi < N_RB * N_GB; wi += NW) {
        const int rb = wi / N_GB, gb = wi - rb * N_GB;
        const int rm = rb * RPW + rsub;
        const int gx = gb * GPW + gxl;
        const float *row = smem + rm * PM + 2 * RX * gx;
        float2 acc[RX];
        analysis_run<RX, J, true, MODE, true>(row, 1, taps.t, a.nch, Fp, nz, acc);
        const int oxb = ox0 + RX * gx;
        if (MODE == kExact && edge_x) {
#pragma unroll
            for (int cc = 0; cc < RX; ++cc) {
                const int i = 2 * (oxb + cc) + 1;
                if (i >= nx && i - nx <= a.F - 2)
                    acc[cc] = analysis_right_edge_p(taps.t + a.shift, a.F, row + 2 * cc + Fp - 1 - a.shift, 1, i - nx);
            }
        }
        const bool low_rows = rm < TR;
        const int oy = oy0 + (low_rows ? rm : rm - TR);
        if (oy < my) {
            const size_t o = (size_t)oy * a.cA.pitch + oxb;
            float *d0 = (low_rows ? pA : pH) + o, *d1 = (low_rows ? pV : pD) + o;
#pragma unroll
            for (int h = 0; h < RX / 4; ++h) {
                if (oxb + 4 * h < mx) {
                    *reinterpret_cast<float4 *>(d0 + 4 * h) =
                        make_float4(acc[4 * h].x, acc[4 * h + 1].x, acc[4 * h + 2].x, acc[4 * h + 3].x);
                    *reinterpret_cast<float4 *>(d1 + 4 * h) =
                        make_float4(acc[4 * h].y, acc[4 * h + 1].y, acc[4 * h + 2].y, acc[4 * h + 3].y);
                }
            }
        }
    }
}

template <int JH, int MODE>
__global__ void __launch_bounds__(kLongNT) k_dwt_inv_x_long(const __grid_constant__ InvTaps taps, const InvLongArgs a)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int TR = kIxTR, TP = kIxTP, RX = kIxR, NT = kLongNT, NW = NT / 32;
    const int Hp = a.Hp;
    const int rp = TP + Hp - 1;
    const int PS = pitch_quads_4mod8(rp);
    const int c4n = (rp + 3) >> 2;
    const int sub_floats = TR * PS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my = a.cH.rows, mx = a.cH.cols;
    const int plane = blockIdx.z, q0 = blockIdx.y * TR, p0 = blockIdx.x * TP;
    const int cx0 = p0 + a.H - Hp + a.shift;
    const bool aligned = (cx0 & 3) == 0;
    const float2 nz = make_float2(a.negzero, a.negzero);
    // [cA, cV, cH, cD][TR][PS]; everything outside the sub-band is zero
    for (int idx = tid; idx < 4 * TR * c4n; idx += NT) {
        const int row = idx / c4n, c4 = idx - row * c4n;
        const int sb = row / TR, ry = row - sb * TR;
        const int y = q0 + ry;
        const B2sImg &im = sb == 0 ? a.cA : (sb == 1 ? a.cV : (sb == 2 ? a.cH : a.cD));
        float *dst = smem + sb * sub_floats + ry * PS + 4 * c4;
        const bool row_ok = y < my;
        const float *srow = im.ptr + (size_t)plane * im.plane_stride + (size_t)(row_ok ? y : 0) * im.pitch;
        const int x = cx0 + 4 * c4;
        if (row_ok && aligned && x >= 0 && x + 3 < mx) {
            cp_async16(dst, srow + x);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (row_ok && x + k >= 0 && x + k < mx) cp_async4(dst + k, srow + x + k);
                else dst[k] = 0.f;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const int gxl = lane & 3, rsub = lane >> 2;
    constexpr int N_GB = TP / RX / 4;
    constexpr int N_RB = (2 * TR) / 8;
    for (int wi = warp; wi < N_RB * N_GB; wi += NW) {
        const int rb = wi / N_GB, gb = wi - rb * N_GB;
        const int rr = rb * 8 + rsub;
        const int sel = rr >= TR ? 1 : 0;
        const int ry = rr - sel * TR;
        const int gx = gb * 4 + gxl;
        const float *rl = smem + (2 * sel) * sub_floats + ry * PS + RX * gx;
        const float *rh = rl + sub_floats;
        float2 o[RX];
        synthesis_run<RX, JH, true, MODE, true>(rl, rh, 1, taps.lo, taps.hi, a.nch, Hp, nz, o);
        const int y = q0 + ry, xo = 2 * (p0 + RX * gx);
        if (y < my) {
            const B2sImg &dimg = sel ? a.d : a.a;
            float *dst = dimg.ptr + (size_t)plane * dimg.plane_stride + (size_t)y * dimg.pitch + xo;
#pragma unroll
            for (int h = 0; h < RX / 2; ++h)
                if (xo + 4 * h < dimg.pitch)
                    *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(o[2 * h].x, o[2 * h].y, o[2 * h + 1].x, o[2 * h + 1].y);
        }
    }
}

template <int JH, int MODE>
__global__ void __launch_bounds__(kLongNT) k_dwt_inv_y_long(const __grid_constant__ InvTaps taps, const InvLongArgs a)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int TQ = kIyTQ, TC = kIyTC, RY = kIyR, NT = kLongNT, NW = NT / 32;
    const int Hp = a.Hp;
    const int rq = TQ + Hp - 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my = a.a.rows;
    const int plane = blockIdx.z, q0 = blockIdx.y * TQ, x0 = blockIdx.x * TC;
    const int cy0 = q0 + a.H - Hp + a.shift;
    const float2 nz = make_float2(a.negzero, a.negzero);
    float *s_a = smem, *s_d = smem + rq * TC;
    const float *ga = a.a.ptr + (size_t)plane * a.a.plane_stride, *gd = a.d.ptr + (size_t)plane * a.d.plane_stride;
    for (int idx = tid; idx < 2 * rq * (TC / 4); idx += NT) {
        const int row = idx / (TC / 4), q = idx % (TC / 4);
        const int sel = row >= rq ? 1 : 0, r = row - sel * rq;
        const int y = cy0 + r, gx = x0 + 4 * q;
        float *dst = (sel ? s_d : s_a) + r * TC + 4 * q;
        if (y >= 0 && y < my && gx < a.a.pitch) cp_async16(dst, (sel ? gd : ga) + (size_t)y * a.a.pitch + gx);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    __syncthreads();
    constexpr int NGQ = TQ / RY, NCG = TC / 32;
    float *dstp = a.out.ptr + (size_t)plane * a.out.plane_stride;
    for (int wi = warp; wi < NGQ * NCG; wi += NW) {
        const int gq = wi / NCG;
        const int x = (wi - gq * NCG) * 32 + lane;
        const float *ca = s_a + (RY * gq) * TC + x;
        const float *cd = s_d + (RY * gq) * TC + x;
        float2 o[RY];
        synthesis_run<RY, JH, true, MODE, false>(ca, cd, TC, taps.lo, taps.hi, a.nch, Hp, nz, o);
        const int ox = x0 + x;
        if (ox < a.out.cols) {
            const int oyb = 2 * (q0 + RY * gq);
#pragma unroll
            for (int k = 0; k < RY; ++k) {
                if (oyb + 2 * k < a.out.rows) dstp[(size_t)(oyb + 2 * k) * a.out.pitch + ox] = o[k].x;
                if (oyb + 2 * k + 1 < a.out.rows) dstp[(size_t)(oyb + 2 * k + 1) * a.out.pitch + ox] = o[k].y;
            }
        }
    }
}

// ---- chunk selection ----------------------------------------------------------------------------------------------
// F <= 20: one chunk of J = F taps (everything compile-time).  Longer filters: chunks of J in {8,12,16,20} taps, the
// choice that pads least (ties: the longer chunk).
void pick_fwd_chunk(int F, int *J, int *nch)
{
    if (F <= 20) { *J = F; *nch = 1; return; }
    int best = 0, best_fp = 1 << 30;
    for (int j : {8, 12, 16, 20}) {
        const int fp = (F + j - 1) / j * j;
        if (fp <= best_fp) { best_fp = fp; best = j; }
    }
    *J = best;
    *nch = best_fp / best;
}
void pick_inv_chunk(int H, int *JH, int *nch)
{
    if (H <= 10) { *JH = H; *nch = 1; return; }
    int best = 0, best_hp = 1 << 30;
    for (int j : {4, 8, 12}) {
        const int hp = (H + j - 1) / j * j;
        if (hp <= best_hp) { best_hp = hp; best = j; }
    }
    *JH = best;
    *nch = best_hp / best;
}

// tile shapes (selectable with B2S_DWT_CFG for tuning): 0 = R 4, 256 threads; 1 = R 8, 128 threads
typedef Tile<16, 64, 256, 4> FwdTile0;
typedef Tile<32, 32, 256, 4> InvTile0;
typedef Tile<32, 64, 256, 4, 8> FwdTile1;   // y-pass 40 warp items, x-pass 16
typedef Tile<31, 64, 256, 8, 4> InvTile1;   // x-pass 20 warp items, y-pass 32
typedef Tile<32, 64, 256, 8, 8> FwdTile2;
// long filters (chunked taps): one CTA per SM fits, so it carries 16 warps to cover the FADD2 dependency chains
typedef Tile<16, 64, 512, 4> FwdTileL;
typedef Tile<32, 32, 512, 4> InvTileL;
typedef Tile<31, 64, 256, 8, 8> InvTile2;

int dev_knob(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <class T, int J, bool MULTI, int MODE>
void launch_fwd_t(const FwdTaps &ft, FwdArgs a, int n_planes, int sm_count, cudaStream_t s)
{
    const FwdGeom g(a.Fp, T::TY, T::TX);
    const size_t bytes = g.smem_bytes() + 128;   // + alignment slack of the TMA destination
    a.tiles_x = (a.cA.cols + T::TX - 1) / T::TX;
    a.tiles_y = (a.cA.rows + T::TY - 1) / T::TY;
    a.n_tiles = a.tiles_x * a.tiles_y * n_planes;
    int per_sm = (int)(kSmemPerSm / (bytes + 1024));
    if (per_sm > 2048 / T::NT) per_sm = 2048 / T::NT;
    if (per_sm < 1) per_sm = 1;
    static const int persist = dev_knob("B2S_DWT_PERSIST", 0);
    const bool one_per_cta = !persist || a.n_tiles < sm_count * per_sm;
    a.grid3d = one_per_cta && n_planes <= 65535 && a.tiles_y <= 65535;
    const dim3 grid = a.grid3d ? dim3(a.tiles_x, a.tiles_y, n_planes) : dim3(one_per_cta ? a.n_tiles : sm_count * per_sm);
    CUtensorMap tmap;
    a.use_tma = make_tmap(&tmap, a.in, n_planes, g.pin, g.rin_y);
    cudaFuncSetAttribute(k_dwt_fwd<T, J, MULTI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_dwt_fwd<T, J, MULTI, MODE><<<grid, T::NT, bytes, s>>>(ft, a, tmap);
}
template <int J, bool MULTI>
void launch_fwd_j(const FwdTaps &ft, const FwdArgs &a, int n_planes, int exact, int sm_count, cudaStream_t s)
{
    static const int cfg = dev_knob("B2S_DWT_CFG", 0);
    const bool fits1 = FwdGeom(a.Fp, FwdTile1::TY, FwdTile1::TX).smem_bytes() <= (size_t)kSmemPerSm - 1024;
    if (cfg == 1 && fits1 && J == 20 && !MULTI) {
        constexpr int JJ = (J == 20 && !MULTI) ? J : 2;   // R = 8 is instantiated for the headline filter length only
        if (exact) launch_fwd_t<FwdTile1, JJ, false, kExact>(ft, a, n_planes, sm_count, s);
        else launch_fwd_t<FwdTile1, JJ, false, kFast>(ft, a, n_planes, sm_count, s);
        return;
    }
    if (cfg == 2 && J == 20 && !MULTI) {
        constexpr int JJ = (J == 20 && !MULTI) ? J : 2;
        if (exact) launch_fwd_t<FwdTile2, JJ, false, kExact>(ft, a, n_planes, sm_count, s);
        else launch_fwd_t<FwdTile2, JJ, false, kFast>(ft, a, n_planes, sm_count, s);
        return;
    }
    if (MULTI) {
        constexpr bool M = MULTI;   // instantiate the wide tile for the chunked variants only
        if (exact) launch_fwd_t<FwdTileL, J, M, kExact>(ft, a, n_planes, sm_count, s);
        else launch_fwd_t<FwdTileL, J, M, kFast>(ft, a, n_planes, sm_count, s);
        return;
    }
    if (exact) launch_fwd_t<FwdTile0, J, MULTI, kExact>(ft, a, n_planes, sm_count, s);
    else launch_fwd_t<FwdTile0, J, MULTI, kFast>(ft, a, n_planes, sm_count, s);
}

template <class T, int JH, bool MULTI, int MODE>
void launch_inv_t(const InvTaps &it, InvArgs a, int n_planes, int sm_count, cudaStream_t s)
{
    const InvGeom g(a.Hp, T::TY, T::TX, T::R1);
    const size_t bytes = g.smem_bytes() + 128;
    a.tiles_x = (a.out.cols + 2 * T::TX - 1) / (2 * T::TX);
    a.tiles_y = (a.out.rows + 2 * T::TY - 1) / (2 * T::TY);
    a.n_tiles = a.tiles_x * a.tiles_y * n_planes;
    int per_sm = (int)(kSmemPerSm / (bytes + 1024));
    if (per_sm > 2048 / T::NT) per_sm = 2048 / T::NT;
    if (per_sm < 1) per_sm = 1;
    static const int persist = dev_knob("B2S_DWT_PERSIST", 0);
    const bool one_per_cta = !persist || a.n_tiles < sm_count * per_sm;
    a.grid3d = one_per_cta && n_planes <= 65535 && a.tiles_y <= 65535;
    const dim3 grid = a.grid3d ? dim3(a.tiles_x, a.tiles_y, n_planes) : dim3(one_per_cta ? a.n_tiles : sm_count * per_sm);
    InvMaps maps;
    // cA is read with the detail bands' logical size (the caller's buffer may be the larger reconstruction target)
    B2sImg ca = a.cA;
    ca.rows = a.cH.rows; ca.cols = a.cH.cols;
    a.use_tma = make_tmap(&maps.m[0], ca, n_planes, g.ps, g.rq) && make_tmap(&maps.m[1], a.cV, n_planes, g.ps, g.rq) &&
                make_tmap(&maps.m[2], a.cH, n_planes, g.ps, g.rq) && make_tmap(&maps.m[3], a.cD, n_planes, g.ps, g.rq);
    cudaFuncSetAttribute(k_dwt_inv<T, JH, MULTI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_dwt_inv<T, JH, MULTI, MODE><<<grid, T::NT, bytes, s>>>(it, a, maps);
}
template <int JH, bool MULTI>
void launch_inv_j(const InvTaps &it, const InvArgs &a, int n_planes, int exact, int sm_count, cudaStream_t s)
{
    static const int cfg = dev_knob("B2S_DWT_CFG", 0);
    if (cfg == 1 && JH == 10 && !MULTI) {
        constexpr int JJ = (JH == 10 && !MULTI) ? JH : 1;
        if (exact) launch_inv_t<InvTile1, JJ, false, kExact>(it, a, n_planes, sm_count, s);
        else launch_inv_t<InvTile1, JJ, false, kFast>(it, a, n_planes, sm_count, s);
        return;
    }
    if (cfg == 2 && JH == 10 && !MULTI) {
        constexpr int JJ = (JH == 10 && !MULTI) ? JH : 1;
        if (exact) launch_inv_t<InvTile2, JJ, false, kExact>(it, a, n_planes, sm_count, s);
        else launch_inv_t<InvTile2, JJ, false, kFast>(it, a, n_planes, sm_count, s);
        return;
    }
    if (MULTI) {
        constexpr bool M = MULTI;
        if (exact) launch_inv_t<InvTileL, JH, M, kExact>(it, a, n_planes, sm_count, s);
        else launch_inv_t<InvTileL, JH, M, kFast>(it, a, n_planes, sm_count, s);
        return;
    }
    if (exact) launch_inv_t<InvTile0, JH, MULTI, kExact>(it, a, n_planes, sm_count, s);
    else launch_inv_t<InvTile0, JH, MULTI, kFast>(it, a, n_planes, sm_count, s);
}


constexpr int kLongMinF = 42;   // filters at least this long take the per-axis kernels when a scratch buffer is given

template <int J>
void launch_fwd_long(const B2sTaps &t, FwdLongArgs a, int nch, int n_planes, int exact, cudaStream_t s)
{
    a.F = t.F; a.Fp = J * nch; a.nch = nch; a.negzero = -0.0f;
    // leading zero taps: the axis -1 windows start at column 2*ox0 + 2 - Fp + shift, which must be a multiple of 4
    a.shift = 0;
    for (int sh = 0; sh <= a.Fp - t.F; ++sh)
        if (((2 - a.Fp + sh) & 3) == 0) { a.shift = sh; break; }
    FwdTaps ft;
    for (int j = 0; j < kMaxFp; ++j) {
        const int k = j - a.shift;
        ft.t[j] = (k >= 0 && k < t.F) ? make_float2(t.dec_lo[k], t.dec_hi[k]) : make_float2(0.f, 0.f);
    }
    const int my = a.cA.rows, mx = a.cA.cols, nx = a.in.cols;
    const size_t smem_y = sizeof(float) * (size_t)(2 * kFyTY + a.Fp - 2) * kFyTC;
    const size_t smem_x = sizeof(float) * (size_t)(2 * kFxTR) * pitch_quads_odd(2 * kFxTX + a.Fp - 2 + 2);
    const dim3 gy((nx + kFyTC - 1) / kFyTC, (my + kFyTY - 1) / kFyTY, n_planes);
    const dim3 gx((mx + kFxTX - 1) / kFxTX, (my + kFxTR - 1) / kFxTR, n_planes);
    if (exact) {
        cudaFuncSetAttribute(k_dwt_fwd_y_long<J, kExact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y);
        cudaFuncSetAttribute(k_dwt_fwd_x_long<J, kExact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
        k_dwt_fwd_y_long<J, kExact><<<gy, kLongNT, smem_y, s>>>(ft, a);
        k_dwt_fwd_x_long<J, kExact><<<gx, kLongNT, smem_x, s>>>(ft, a);
    } else {
        cudaFuncSetAttribute(k_dwt_fwd_y_long<J, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y);
        cudaFuncSetAttribute(k_dwt_fwd_x_long<J, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
        k_dwt_fwd_y_long<J, kFast><<<gy, kLongNT, smem_y, s>>>(ft, a);
        k_dwt_fwd_x_long<J, kFast><<<gx, kLongNT, smem_x, s>>>(ft, a);
    }
}

template <int JH>
void launch_inv_long(const B2sTaps &t, InvLongArgs a, int nch, int n_planes, int exact, cudaStream_t s)
{
    const int H = t.F / 2;
    a.H = H; a.Hp = JH * nch; a.nch = nch; a.negzero = -0.0f;
    a.shift = a.Hp - H;   // all padding in front: the coefficient windows start at the tile origin
    InvTaps it;
    for (int j = 0; j < kMaxFp / 2; ++j) {
        const int k = j - a.shift;
        const bool ok = k >= 0 && k < H;
        it.lo[j] = ok ? make_float2(t.rec_lo[2 * k], t.rec_lo[2 * k + 1]) : make_float2(0.f, 0.f);
        it.hi[j] = ok ? make_float2(t.rec_hi[2 * k], t.rec_hi[2 * k + 1]) : make_float2(0.f, 0.f);
    }
    const int my = a.cH.rows;
    const size_t smem_x = sizeof(float) * 4 * (size_t)kIxTR * pitch_quads_4mod8(kIxTP + a.Hp - 1);
    const size_t smem_y = sizeof(float) * 2 * (size_t)(kIyTQ + a.Hp - 1) * kIyTC;
    const dim3 gx((a.out.cols + 2 * kIxTP - 1) / (2 * kIxTP), (my + kIxTR - 1) / kIxTR, n_planes);
    const dim3 gy((a.out.cols + kIyTC - 1) / kIyTC, (a.out.rows + 2 * kIyTQ - 1) / (2 * kIyTQ), n_planes);
    if (exact) {
        cudaFuncSetAttribute(k_dwt_inv_x_long<JH, kExact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
        cudaFuncSetAttribute(k_dwt_inv_y_long<JH, kExact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y);
        k_dwt_inv_x_long<JH, kExact><<<gx, kLongNT, smem_x, s>>>(it, a);
        k_dwt_inv_y_long<JH, kExact><<<gy, kLongNT, smem_y, s>>>(it, a);
    } else {
        cudaFuncSetAttribute(k_dwt_inv_x_long<JH, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
        cudaFuncSetAttribute(k_dwt_inv_y_long<JH, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y);
        k_dwt_inv_x_long<JH, kFast><<<gx, kLongNT, smem_x, s>>>(it, a);
        k_dwt_inv_y_long<JH, kFast><<<gy, kLongNT, smem_y, s>>>(it, a);
    }
}

B2sImg scratch_img(float *base, size_t plane_stride, int half, int rows, int cols, int pitch)
{
    B2sImg im;
    im.ptr = base + (size_t)half * rows * pitch;
    im.plane_stride = plane_stride;
    im.pitch = pitch;
    im.rows = rows;
    im.cols = cols;
    return im;
}

}  // namespace

int b2s_dwt_max_smem(int F)
{
    int J, nch, JH, nchi;
    pick_fwd_chunk(F, &J, &nch);
    pick_inv_chunk(F / 2, &JH, &nchi);
    const size_t a = FwdGeom(J * nch, FwdTile0::TY, FwdTile0::TX).smem_bytes();
    const size_t b = InvGeom(JH * nchi, InvTile0::TY, InvTile0::TX, InvTile0::R1).smem_bytes();
    return (int)((a > b ? a : b) + 128 + 1024);
}

size_t b2s_dwt_scratch_floats(int F, int ny, int nx)
{
    static const int off = dev_knob("B2S_DWT_NO_LONG", 0);
    if (F < kLongMinF || off) return 0;
    return 2 * (size_t)((ny + F - 1) / 2) * round_up4(nx);
}

void b2s_launch_dwt_fwd(const B2sTaps &t, const B2sImg &in, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV,
                        const B2sImg &cD, int n_planes, int exact, int sm_count, cudaStream_t s, float *scratch,
                        size_t scratch_plane_stride)
{
    int J, nch;
    pick_fwd_chunk(t.F, &J, &nch);
    if (scratch && t.F >= kLongMinF) {
        FwdLongArgs a;
        a.in = in; a.cA = cA; a.cH = cH; a.cV = cV; a.cD = cD;
        const int pitch = round_up4(in.cols);
        a.lo = scratch_img(scratch, scratch_plane_stride, 0, cA.rows, in.cols, pitch);
        a.hi = scratch_img(scratch, scratch_plane_stride, 1, cA.rows, in.cols, pitch);
        switch (J) {
#define X(JJ) case JJ: launch_fwd_long<JJ>(t, a, nch, n_planes, exact, s); return;
            X(8) X(12) X(16) X(20)
#undef X
        }
    }
    FwdTaps ft;
    for (int j = 0; j < kMaxFp; ++j) ft.t[j] = j < t.F ? make_float2(t.dec_lo[j], t.dec_hi[j]) : make_float2(0.f, 0.f);
    FwdArgs a;
    a.in = in; a.cA = cA; a.cH = cH; a.cV = cV; a.cD = cD;
    a.F = t.F; a.Fp = J * nch; a.nch = nch;
    a.negzero = -0.0f;
    if (nch == 1) {
        switch (J) {
#define X(JJ) case JJ: launch_fwd_j<JJ, false>(ft, a, n_planes, exact, sm_count, s); return;
            X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20)
#undef X
        }
    }
    switch (J) {
#define X(JJ) case JJ: launch_fwd_j<JJ, true>(ft, a, n_planes, exact, sm_count, s); return;
        X(8) X(12) X(16) X(20)
#undef X
    }
}

void b2s_launch_dwt_inv(const B2sTaps &t, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV, const B2sImg &cD,
                        const B2sImg &out, int n_planes, int exact, int sm_count, cudaStream_t s, float *scratch,
                        size_t scratch_plane_stride)
{
    const int H = t.F / 2;
    int JH, nch;
    pick_inv_chunk(H, &JH, &nch);
    if (scratch && t.F >= kLongMinF) {
        InvLongArgs a;
        a.cA = cA; a.cH = cH; a.cV = cV; a.cD = cD; a.out = out;
        const int pitch = round_up4(out.cols);
        a.a = scratch_img(scratch, scratch_plane_stride, 0, cH.rows, out.cols, pitch);
        a.d = scratch_img(scratch, scratch_plane_stride, 1, cH.rows, out.cols, pitch);
        switch (JH) {
#define X(JJ) case JJ: launch_inv_long<JJ>(t, a, nch, n_planes, exact, s); return;
            X(4) X(8) X(12)
#undef X
        }
    }
    InvTaps it;
    for (int j = 0; j < kMaxFp / 2; ++j) {
        it.lo[j] = j < H ? make_float2(t.rec_lo[2 * j], t.rec_lo[2 * j + 1]) : make_float2(0.f, 0.f);
        it.hi[j] = j < H ? make_float2(t.rec_hi[2 * j], t.rec_hi[2 * j + 1]) : make_float2(0.f, 0.f);
    }
    InvArgs a;
    a.cA = cA; a.cH = cH; a.cV = cV; a.cD = cD; a.out = out;
    a.H = H; a.Hp = JH * nch; a.nch = nch;
    a.negzero = -0.0f;
    if (nch == 1) {
        switch (JH) {
#define X(JJ) case JJ: launch_inv_j<JJ, false>(it, a, n_planes, exact, sm_count, s); return;
            X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10)
#undef X
        }
    }
    switch (JH) {
#define X(JJ) case JJ: launch_inv_j<JJ, true>(it, a, n_planes, exact, sm_count, s); return;
        X(4) X(8) X(12)
#undef X
    }
}
