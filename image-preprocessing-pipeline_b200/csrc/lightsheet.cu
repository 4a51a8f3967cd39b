// lightsheet_correct on the GPU: local percentiles on a sub-grid, scipy.ndimage.zoom(order=1) back to the image, and
// the subtraction  img -= min(img, min(ls, bg * weight)).
//
// Replaces pystripe/lightsheet_correct.py:31-106 (correct_lightsheet), :113-237 (apply_local_function),
// :240-242 (prctl = numba np.percentile) as called from pystripe/core.py:1333-1348.
//   * window percentile: numba's formula  rank = 1 + (n-1) q/100, f = floor(rank), m = rank - f,
//     v = a[f-1] (1-m) + a[min(f, n-1)] m  in float64, stored into a d_type (uint16) grid with a C cast (truncation).
//     The two order statistics are found by a bit-serial binary search over order-preserving integer keys (16 probes for
//     uint16 pixels, 32 for float32): every thread keeps its share of the window in registers and counts keys below the
//     probe; no sort, no histogram, exact for any data.
//   * zoom: scipy's NI_ZoomShift for order 1, mode 'constant', grid_mode False, restated operation for operation in
//     float64 (coordinate o * (in-1)/(out-1); a coordinate that rounds above in-1 yields 0 exactly as scipy does;
//     weights w0 = 1 - x, w1 = 1 - w0; value ((g*wy)*wx) summed over the 2x2 footprint in row-major order; integer
//     output (t + 0.5) truncated).  Per-axis tables (indices, weights, zero flags) are built once per plan.
//   This file is compiled with -fmad=false: no double-precision contraction.
#include <cstdlib>

#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned load_key(const void *p, int dtype, size_t idx)
{
    if (dtype == B2S_U16) return __ldg(reinterpret_cast<const unsigned short *>(p) + idx);
    if (dtype == B2S_U8) return __ldg(reinterpret_cast<const unsigned char *>(p) + idx);
    return f2key(__ldg(reinterpret_cast<const float *>(p) + idx));
}

// sum over the NT threads that work on one window (NT == 32: one warp; otherwise the whole CTA)
template <int NT>
__device__ __forceinline__ unsigned group_sum(unsigned v, unsigned *scratch)
{
    v = __reduce_add_sync(0xffffffffu, v);
    if (NT == 32) return v;
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[warp] = v;
    __syncthreads();
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += scratch[w];
    return s;
}
template <int NT>
__device__ __forceinline__ unsigned group_min(unsigned v, unsigned *scratch)
{
    v = __reduce_min_sync(0xffffffffu, v);
    if (NT == 32) return v;
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[warp] = v;
    __syncthreads();
    unsigned s = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s = min(s, scratch[w]);
    return s;
}

struct AxisGeom { int n, left, spacing, hl, hr, step, size; };   // centres c = left + i*spacing, window [c-hl, c+hr) clipped

struct GridArgs {
    const void *img;       // post-dark image, (rows x cols) per plane
    int dtype;
    int rows, cols;
    AxisGeom gy, gx;
    double qfrac;          // q / 100
    int q_is_100;
    unsigned short *grid;  // (gy.n x gx.n) per plane
};

// NT threads per window; KMAX keys per thread in registers
template <int NT, int KMAX>
__global__ void __launch_bounds__(NT == 32 ? 256 : NT) k_window_percentile(const GridArgs a)
{
    __shared__ unsigned scratch[32];
    const int lane_in_group = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int groups_per_cta = NT == 32 ? 8 : 1;
    const int win = blockIdx.x * groups_per_cta + (NT == 32 ? (threadIdx.x >> 5) : 0);
    const int n_win = a.gy.n * a.gx.n;
    if (NT == 32 && win >= n_win) return;     // whole warps only
    const int iy = win / a.gx.n, ix = win - iy * a.gx.n;
    const int cy = a.gy.left + iy * a.gy.spacing, cx = a.gx.left + ix * a.gx.spacing;
    const int y0 = max(0, cy - a.gy.hl), y1 = min(cy + a.gy.hr, a.gy.size);
    const int x0 = max(0, cx - a.gx.hl), x1 = min(cx + a.gx.hr, a.gx.size);
    const int ny = (y1 - y0 + a.gy.step - 1) / a.gy.step, nx = (x1 - x0 + a.gx.step - 1) / a.gx.step;
    const int n = ny * nx;
    const char *plane = reinterpret_cast<const char *>(a.img) +
                        (size_t)blockIdx.y * a.rows * a.cols * (a.dtype == B2S_F32 ? 4 : (a.dtype == B2S_U16 ? 2 : 1));
    unsigned key[KMAX];
    int mine = 0;
#pragma unroll
    for (int s = 0; s < KMAX; ++s) {
        const int t = lane_in_group + s * NT;
        key[s] = 0xffffffffu;
        if (t < n) {
            const int ty = t / nx, tx = t - ty * nx;
            key[s] = load_key(plane, a.dtype, (size_t)(y0 + ty * a.gy.step) * a.cols + (x0 + tx * a.gx.step));
            mine = s + 1;
        }
    }
    // numba np.percentile ranks
    double value = 0.0;
    if (n > 0) {
        int k0, k1;
        double m = 0.0;
        if (n == 1) { k0 = k1 = 0; }
        else if (a.q_is_100) { k0 = k1 = n - 1; }
        else {
            const double rank = 1.0 + (double)(n - 1) * a.qfrac;
            const double f = floor(rank);
            m = rank - f;
            k0 = (int)f - 1;
            k1 = min((int)f, n - 1);
        }
        // k0-th smallest key: greatest t such that fewer than k0+1 keys are below t
        const int top = a.dtype == B2S_F32 ? 31 : (a.dtype == B2S_U16 ? 15 : 7);
        unsigned t = 0;
        for (int b = top; b >= 0; --b) {
            const unsigned p = t | (1u << b);
            unsigned c = 0;
#pragma unroll
            for (int s = 0; s < KMAX; ++s) c += (s < mine && key[s] < p) ? 1u : 0u;
            c = group_sum<NT>(c, scratch);
            if (c <= (unsigned)k0) t = p;
        }
        unsigned v0 = t, v1 = t;
        if (k1 != k0) {
            unsigned le = 0, nxt = 0xffffffffu;
#pragma unroll
            for (int s = 0; s < KMAX; ++s) {
                if (s < mine) {
                    le += key[s] <= v0 ? 1u : 0u;
                    if (key[s] > v0) nxt = min(nxt, key[s]);
                }
            }
            le = group_sum<NT>(le, scratch);
            nxt = group_min<NT>(nxt, scratch);
            if (le <= (unsigned)k1) v1 = nxt;
        }
        const double lower = a.dtype == B2S_F32 ? (double)key2f(v0) : (double)v0;
        const double upper = a.dtype == B2S_F32 ? (double)key2f(v1) : (double)v1;
        value = (n == 1 || a.q_is_100) ? lower : lower * (1.0 - m) + upper * m;
    }
    if (lane_in_group == 0)
        a.grid[(size_t)blockIdx.y * n_win + win] = (unsigned short)(long long)value;   // C cast: truncation
}

// Integer pixels (uint8 / uint16), one CTA per window: two 16-bit keys per register, compared two at a time with the SIMD
// video instructions (__vsetltu2 yields 0/1 per half-word), sample coordinates advanced incrementally (no division per
// sample), the search started at the highest bit set in the window's maximum.  Missing samples are padded with 0xffff,
// which is never below a probe.
template <int NT, int KP>   // KP packed registers per thread: up to 2*KP*NT samples
__global__ void __launch_bounds__(NT) k_window_percentile_u16(const GridArgs a)
{
    __shared__ unsigned scratch[32];
    const int win = blockIdx.x;
    const int n_win = a.gy.n * a.gx.n;
    const int iy = win / a.gx.n, ix = win - iy * a.gx.n;
    const int cy = a.gy.left + iy * a.gy.spacing, cx = a.gx.left + ix * a.gx.spacing;
    const int y0 = max(0, cy - a.gy.hl), y1 = min(cy + a.gy.hr, a.gy.size);
    const int x0 = max(0, cx - a.gx.hl), x1 = min(cx + a.gx.hr, a.gx.size);
    const int ny = (y1 - y0 + a.gy.step - 1) / a.gy.step, nx = (x1 - x0 + a.gx.step - 1) / a.gx.step;
    const int n = ny * nx;
    const size_t plane_off = (size_t)blockIdx.y * a.rows * a.cols;
    const unsigned short *p16 = reinterpret_cast<const unsigned short *>(a.img) + plane_off;
    const unsigned char *p8 = reinterpret_cast<const unsigned char *>(a.img) + plane_off;
    // sample t = tid + s*NT -> (ty, tx); advancing t by NT moves (ty, tx) by (dq, dr) with carry
    const int dq = NT / nx, dr = NT - dq * nx;
    int ty = threadIdx.x / nx, tx = threadIdx.x - ty * nx;
    unsigned key[KP];
    unsigned vmax = 0;
#pragma unroll
    for (int s = 0; s < KP; ++s) {
        unsigned w = 0;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            unsigned v = 0xffffu;
            if (ty < ny) {
                const size_t idx = (size_t)(y0 + ty * a.gy.step) * a.cols + (x0 + tx * a.gx.step);
                v = a.dtype == B2S_U16 ? __ldg(p16 + idx) : __ldg(p8 + idx);
                vmax = max(vmax, v);
            }
            w |= v << (16 * hlf);
            ty += dq; tx += dr;
            if (tx >= nx) { tx -= nx; ++ty; }
        }
        key[s] = w;
    }
    double value = 0.0;
    if (n > 0) {
        int k0, k1;
        double m = 0.0;
        if (n == 1) { k0 = k1 = 0; }
        else if (a.q_is_100) { k0 = k1 = n - 1; }
        else {
            const double rank = 1.0 + (double)(n - 1) * a.qfrac;
            const double f = floor(rank);
            m = rank - f;
            k0 = (int)f - 1;
            k1 = min((int)f, n - 1);
        }
        // window maximum bounds the search
        vmax = __reduce_max_sync(0xffffffffu, vmax);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = vmax;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) vmax = max(vmax, scratch[w]);
        const int top = vmax ? 31 - __clz(vmax) : 0;
        unsigned t = 0;
        for (int b = top; b >= 0; --b) {
            const unsigned p = t | (1u << b);
            const unsigned pp = p | (p << 16);
            unsigned c2 = 0;
#pragma unroll
            for (int s = 0; s < KP; ++s) c2 += __vsetltu2(key[s], pp);
            unsigned c = (c2 & 0xffffu) + (c2 >> 16);
            c = group_sum<NT>(c, scratch);
            if (c <= (unsigned)k0) t = p;
        }
        unsigned v0 = t, v1 = t;
        if (k1 != k0) {
            // keys <= v0 and the smallest key above v0
            unsigned le;
            if (v0 >= 0xffffu) le = (unsigned)n;
            else {
                const unsigned q = v0 + 1, qq = q | (q << 16);
                unsigned c2 = 0;
#pragma unroll
                for (int s = 0; s < KP; ++s) c2 += __vsetltu2(key[s], qq);
                le = group_sum<NT>((c2 & 0xffffu) + (c2 >> 16), scratch);
            }
            if (le <= (unsigned)k1) {
                unsigned nxt = 0xffffffffu;
                int cnt = 0;   // only real samples: padding is 0xffff and may coincide with a real 0xffff, which is fine (same value)
#pragma unroll
                for (int s = 0; s < KP; ++s) {
                    const unsigned lo = key[s] & 0xffffu, hi = key[s] >> 16;
                    if (lo > v0) nxt = min(nxt, lo);
                    if (hi > v0) nxt = min(nxt, hi);
                }
                (void)cnt;
                v1 = group_min<NT>(nxt, scratch);
            }
        }
        const double lower = (double)v0, upper = (double)v1;
        value = (n == 1 || a.q_is_100) ? lower : lower * (1.0 - m) + upper * m;
    }
    if (threadIdx.x == 0) a.grid[(size_t)blockIdx.y * n_win + win] = (unsigned short)(long long)value;
}

// ---- background windows through sorted cells ---------------------------------------------------------------------
// When the window is a whole number of grid spacings (200 = 8 * 25, the default) every window is a union of S x S cells
// whose boundaries sit at (left - hl) mod S, and the step-2 sub-sampling of a window selects, inside each of its cells,
// the pixels of one parity class (py, px) = parity of the window's (clipped) start.  So: sort every (cell, class) list once
// (each pixel belongs to exactly one list), then answer "how many window samples are below p" as a sum of binary
// searches over the window's cells.  Work per plane drops from 6561 x 10000 x 13 comparisons to one sort of the image
// plus 6561 x 13 x 64 binary searches.  Exactly the same order statistics come out.
constexpr int kCellCap = 176;   // >= 13 * 13 samples of one class in a 25 x 25 cell, multiple of 8

struct CellAxis { int o, S, nc, size; };   // cell j covers [o + S*(j-1), o + S*j) clipped to [0, size)
__device__ __forceinline__ int cell_lo(const CellAxis &a, int j) { return max(0, a.o + a.S * (j - 1)); }
__device__ __forceinline__ int cell_hi(const CellAxis &a, int j) { return min(a.size, a.o + a.S * j); }
__device__ __forceinline__ int class_count(int lo, int hi, int par)   // integers x in [lo, hi) with x % 2 == par
{
    const int first = lo + ((lo & 1) != par);
    return first < hi ? (hi - first + 1) / 2 : 0;
}

struct CellArgs {
    const void *img; int dtype, rows, cols;
    CellAxis cy, cx;
    unsigned short *lists;      // [plane][class = 2*py+px][jy][jx][kCellCap], ascending
};

// one warp per (class, cell): bitonic sort of 256 keys (8 per lane, key e = 8*lane + r) held in registers; partners at
// distance >= 8 are exchanged with warp shuffles, closer ones are register pairs
template <int K, int J>
__device__ __forceinline__ void bitonic_step(unsigned (&v)[8], int lane)
{
    if (J >= 8) {
        constexpr int M = J >> 3;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const unsigned other = __shfl_xor_sync(0xffffffffu, v[r], M);
            const bool up = (((lane << 3) | r) & K) == 0;
            const bool lower = (lane & M) == 0;
            v[r] = (lower == up) ? min(v[r], other) : max(v[r], other);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if ((r & J) == 0) {
                const unsigned x = v[r], y = v[r | J];
                const bool up = (((lane << 3) | r) & K) == 0;
                v[r] = up ? min(x, y) : max(x, y);
                v[r | J] = up ? max(x, y) : min(x, y);
            }
        }
    }
}
template <int K, int J>
struct BitonicMerge {
    static __device__ __forceinline__ void run(unsigned (&v)[8], int lane)
    {
        bitonic_step<K, J>(v, lane);
        BitonicMerge<K, (J >> 1)>::run(v, lane);
    }
};
template <int K>
struct BitonicMerge<K, 0> { static __device__ __forceinline__ void run(unsigned (&)[8], int) {} };
template <int K>
struct BitonicSort {
    static __device__ __forceinline__ void run(unsigned (&v)[8], int lane)
    {
        BitonicSort<(K >> 1)>::run(v, lane);
        BitonicMerge<K, (K >> 1)>::run(v, lane);
    }
};
template <>
struct BitonicSort<1> { static __device__ __forceinline__ void run(unsigned (&)[8], int) {} };

__global__ void __launch_bounds__(256) k_sort_cells(const CellArgs a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cells = a.cy.nc * a.cx.nc;
    const int item = blockIdx.x * 8 + warp;
    if (item >= 4 * n_cells) return;
    const int cls = item / n_cells, cell = item - cls * n_cells;
    const int jy = cell / a.cx.nc, jx = cell - jy * a.cx.nc;
    const int py = cls >> 1, px = cls & 1;
    const int ylo = cell_lo(a.cy, jy), yhi = cell_hi(a.cy, jy), xlo = cell_lo(a.cx, jx), xhi = cell_hi(a.cx, jx);
    const int ny = class_count(ylo, yhi, py), nx = class_count(xlo, xhi, px);
    const int n = ny * nx;
    const int y0 = ylo + ((ylo & 1) != py), x0 = xlo + ((xlo & 1) != px);
    const size_t plane_off = (size_t)blockIdx.y * a.rows * a.cols;
    unsigned v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int e = (lane << 3) | r;
        v[r] = 0xffffu;
        if (e < n) {
            const int ey = e / nx, ex = e - ey * nx;
            const size_t idx = plane_off + (size_t)(y0 + 2 * ey) * a.cols + (x0 + 2 * ex);
            v[r] = a.dtype == B2S_U16 ? __ldg(reinterpret_cast<const unsigned short *>(a.img) + idx)
                                      : __ldg(reinterpret_cast<const unsigned char *>(a.img) + idx);
        }
    }
    BitonicSort<256>::run(v, lane);
    if ((lane << 3) < kCellCap) {
        unsigned short *dst = a.lists + (((size_t)blockIdx.y * 4 + cls) * n_cells + cell) * kCellCap + (lane << 3);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    }
}

struct CellQueryArgs {
    CellAxis cy, cx;
    AxisGeom gy, gx;            // window geometry (centres, half widths)
    const unsigned short *lists;
    double qfrac; int q_is_100;
    unsigned short *grid;
};

// number of keys < p in an ascending list, given that the answer lies in [lo, hi]
__device__ __forceinline__ int lower_bound_u16(const unsigned short *l, int lo, int hi, unsigned p)
{
    int len = hi - lo;
    while (len > 0) {
        const int half = len >> 1;
        if (__ldg(l + lo + half) < p) { lo += half + 1; len -= half + 1; }
        else len = half;
    }
    return lo;
}

// one warp per window; each lane owns up to CPL of the window's cells
template <int CPL>
__global__ void __launch_bounds__(256) k_window_percentile_cells(const CellQueryArgs a)
{
    const int lane = threadIdx.x & 31;
    const int win = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n_win = a.gy.n * a.gx.n;
    if (win >= n_win) return;
    const int iy = win / a.gx.n, ix = win - iy * a.gx.n;
    const int sy_raw = a.gy.left + iy * a.gy.spacing - a.gy.hl, sx_raw = a.gx.left + ix * a.gx.spacing - a.gx.hl;
    const int py = max(0, sy_raw) & 1, px = max(0, sx_raw) & 1;
    const int wy = (a.gy.hl + a.gy.hr) / a.cy.S, wx = (a.gx.hl + a.gx.hr) / a.cx.S;     // cells per window side
    // first cell of the (unclipped) window: sy_raw = o + S*(j-1)
    const int jy0 = (sy_raw - a.cy.o) / a.cy.S + 1, jx0 = (sx_raw - a.cx.o) / a.cx.S + 1;   // exact divisions (may be <= 0)
    const int jylo = max(jy0, 0), jyhi = min(jy0 + wy, a.cy.nc), jxlo = max(jx0, 0), jxhi = min(jx0 + wx, a.cx.nc);
    const int ncx = jxhi - jxlo, ncell = (jyhi - jylo) * ncx;
    const int n_cells = a.cy.nc * a.cx.nc;
    const unsigned short *base = a.lists + ((size_t)blockIdx.y * 4 + (2 * py + px)) * n_cells * kCellCap;
    const unsigned short *lp[CPL];
    int cnt[CPL];
    int n = 0;
    unsigned vmax = 0;
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        const int c = lane + 32 * s;
        cnt[s] = 0;
        lp[s] = base;
        if (c < ncell) {
            const int jy = jylo + c / ncx, jx = jxlo + c % ncx;
            cnt[s] = class_count(cell_lo(a.cy, jy), cell_hi(a.cy, jy), py) * class_count(cell_lo(a.cx, jx), cell_hi(a.cx, jx), px);
            lp[s] = base + ((size_t)jy * a.cx.nc + jx) * kCellCap;
            if (cnt[s] > 0) vmax = max(vmax, (unsigned)__ldg(lp[s] + cnt[s] - 1));
        }
        n += cnt[s];
    }
    n = __reduce_add_sync(0xffffffffu, n);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    double value = 0.0;
    if (n > 0) {
        int k0, k1;
        double m = 0.0;
        if (n == 1) { k0 = k1 = 0; }
        else if (a.q_is_100) { k0 = k1 = n - 1; }
        else {
            const double rank = 1.0 + (double)(n - 1) * a.qfrac;
            const double f = floor(rank);
            m = rank - f;
            k0 = (int)f - 1;
            k1 = min((int)f, n - 1);
        }
        const int top = vmax ? 31 - __clz(vmax) : 0;
        // the answer stays in [t, t + 2^(b+1)); per cell the positions of those two bounds bracket every later search
        int blo[CPL], bhi[CPL];
#pragma unroll
        for (int s = 0; s < CPL; ++s) { blo[s] = 0; bhi[s] = cnt[s]; }
        unsigned t = 0;
        for (int b = top; b >= 0; --b) {
            const unsigned p = t | (1u << b);
            int pos[CPL];
            int c = 0;
#pragma unroll
            for (int s = 0; s < CPL; ++s) { pos[s] = lower_bound_u16(lp[s], blo[s], bhi[s], p); c += pos[s]; }
            c = __reduce_add_sync(0xffffffffu, c);
            const bool take = c <= k0;
            if (take) t = p;
#pragma unroll
            for (int s = 0; s < CPL; ++s) { if (take) blo[s] = pos[s]; else bhi[s] = pos[s]; }
        }
        unsigned v0 = t, v1 = t;
        if (k1 != k0) {
            int le = 0;
            unsigned nxt = 0xffffffffu;
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const int u = lower_bound_u16(lp[s], blo[s], cnt[s], v0 + 1);      // keys <= v0
                le += u;
                if (u < cnt[s]) nxt = min(nxt, (unsigned)__ldg(lp[s] + u));
            }
            le = __reduce_add_sync(0xffffffffu, le);
            nxt = __reduce_min_sync(0xffffffffu, nxt);
            if (le <= k1) v1 = nxt;
        }
        const double lower = (double)v0, upper = (double)v1;
        value = (n == 1 || a.q_is_100) ? lower : lower * (1.0 - m) + upper * m;
    }
    if (lane == 0) a.grid[(size_t)blockIdx.y * n_win + win] = (unsigned short)(long long)value;
}

CellAxis cell_axis(const AxisGeom &g)
{
    CellAxis c;
    c.S = g.spacing;
    c.size = g.size;
    const int start0 = g.left - g.hl;                       // (unclipped) start of window 0
    c.o = ((start0 % c.S) + c.S) % c.S;                     // cell boundaries at o + S*j
    c.nc = (c.size - 1 - c.o) / c.S + 2;                    // cell 0 = [o - S, o) (empty when o == 0)
    return c;
}

// ---- scipy.ndimage.zoom axis tables -------------------------------------------------------------------------------
struct ZoomAxis { const int *i0, *i1; const double *w0, *w1; const unsigned char *zero; };

__global__ void k_zoom_axis(int n_in, int n_out, int *i0, int *i1, double *w0, double *w1, unsigned char *zero)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    // zoom = (in - 1) / (out - 1), 1 where out == 1 (scipy/ndimage/_interpolation.py zoom)
    const double z = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 1.0;
    double cc = (double)o * z;
    const bool out = cc < 0.0 || cc > (double)(n_in - 1);     // map_coordinate, NI_EXTEND_CONSTANT
    zero[o] = out ? 1 : 0;
    if (out) cc = 0.0;
    const double fl = floor(cc);
    const long long start = (long long)fl;
    const double x = cc - fl;
    const double a = 1.0 - x;
    w0[o] = a;
    w1[o] = 1.0 - a;
    long long idx[2] = {start, start + 1};
    for (int l = 0; l < 2; ++l) {
        long long v = idx[l];
        if (n_in <= 1) v = 0;
        else {
            const long long s2 = 2LL * n_in - 2;
            if (v < 0) { v = s2 * (-v / s2) + v; v = v <= 1 - n_in ? v + s2 : -v; }
            else if (v >= n_in) { v -= s2 * (v / s2); if (v >= n_in) v = s2 - v; }
        }
        idx[l] = v;
    }
    i0[o] = (int)idx[0];
    i1[o] = (int)idx[1];
}

__device__ __forceinline__ unsigned zoom_u16(const unsigned short *g, int gcols, const ZoomAxis &zy, const ZoomAxis &zx,
                                             int y, int x)
{
    if (zy.zero[y] || zx.zero[x]) return 0u;
    const int y0 = zy.i0[y], y1 = zy.i1[y], x0 = zx.i0[x], x1 = zx.i1[x];
    const double wy0 = zy.w0[y], wy1 = zy.w1[y], wx0 = zx.w0[x], wx1 = zx.w1[x];
    double t = ((double)g[(size_t)y0 * gcols + x0] * wy0) * wx0;
    t = t + ((double)g[(size_t)y0 * gcols + x1] * wy0) * wx1;
    t = t + ((double)g[(size_t)y1 * gcols + x0] * wy1) * wx0;
    t = t + ((double)g[(size_t)y1 * gcols + x1] * wy1) * wx1;
    t = t > 0.0 ? t + 0.5 : 0.0;
    if (t > 65535.0) t = 65535.0;
    return (unsigned)t;
}

struct FinalArgs {
    const void *img;      // post-dark image
    int dtype;            // U8/U16 (integer path) or F32
    int rows, cols;
    const unsigned short *ls, *bg;
    int ls_cols, bg_cols, ls_size, bg_size;   // grid widths and per-plane sizes
    ZoomAxis ls_y, ls_x, bg_y, bg_x;
    double weight;
    int weight_is_int;    // isinstance(w, float) and integral handling: see below
    int final_mode, shift, out_dtype, flip, rot;
    const unsigned *uniform_mm;
    void *out;
    int out_rows, out_cols;
};

// img - min(img, min(ls, bg * weight)) (lightsheet_correct.py:89-100) and the final conversion of the value (core.py:1361-1369,
// 397-423; dark was applied before the lightsheet step)
__device__ __forceinline__ void final_store(const FinalArgs &a, size_t p, size_t oidx, unsigned ls, unsigned bg, bool zero_plane)
{
    double v = 0.0;
    if (!zero_plane) {
        if (a.dtype != B2S_F32) {
            const unsigned px = a.dtype == B2S_U16 ? reinterpret_cast<const unsigned short *>(a.img)[p]
                                                   : reinterpret_cast<const unsigned char *>(a.img)[p];
            unsigned sub;
            if (a.weight_is_int) {
                // all-uint fast path (lightsheet_correct.py:89-93): bg * int(w) wraps in the array dtype
                const unsigned mask = a.dtype == B2S_U16 ? 0xffffu : 0xffu;
                sub = min(px, min(ls, (bg * (unsigned)(long long)a.weight) & mask));
            } else {
                const double m = fmin((double)px, fmin((double)ls, (double)bg * a.weight));
                sub = (unsigned)m;                                   // .astype(img.dtype)
            }
            v = (double)(px - sub);
        } else {
            const float px = reinterpret_cast<const float *>(a.img)[p];
            const double m = fmin((double)px, fmin((double)ls, (double)bg * a.weight));
            v = (double)__fsub_rn(px, (float)m);                     // float32 array -= float32
        }
    }
    if (a.final_mode == 3) { reinterpret_cast<float *>(a.out)[oidx] = (float)v; return; }
    unsigned u;
    if (a.final_mode == 2) {
        const double c = v < 0.0 ? 0.0 : (v > 65535.0 ? 65535.0 : v);
        u = (unsigned)c;
        const unsigned lower = 1u << a.shift;
        u = (u > 0 && u < lower) ? 1u : (u >> a.shift);
        u = u > 255u ? 255u : u;
    } else {
        const double hi = (a.out_dtype == B2S_U8) ? 255.0 : 65535.0;
        const double c = v < 0.0 ? 0.0 : (v > hi ? hi : v);
        u = (unsigned)c;
    }
    if (a.out_dtype == B2S_U8) reinterpret_cast<unsigned char *>(a.out)[oidx] = (unsigned char)u;
    else reinterpret_cast<unsigned short *>(a.out)[oidx] = (unsigned short)u;
}

// general form: one thread per output pixel, both zooms evaluated from the grids (any rotation)
__global__ void __launch_bounds__(256) k_lightsheet_final(const FinalArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= a.out_cols) return;
    const size_t plane = blockIdx.z;
    const size_t oidx = plane * (size_t)a.out_rows * a.out_cols + (size_t)i * a.out_cols + j;
    const bool zero_plane = a.uniform_mm && a.uniform_mm[2 * plane] == ~a.uniform_mm[2 * plane + 1];
    const int R = a.rows, C = a.cols;
    int y, x;
    switch (a.rot) {
    case 1: y = j; x = C - 1 - i; break;
    case 2: y = R - 1 - i; x = C - 1 - j; break;
    case 3: y = R - 1 - j; x = i; break;
    default: y = i; x = j; break;
    }
    if (a.flip) y = R - 1 - y;
    const size_t p = plane * (size_t)R * C + (size_t)y * C + x;
    unsigned ls = 0, bg = 0;
    if (!zero_plane) {
        ls = zoom_u16(a.ls + plane * a.ls_size, a.ls_cols, a.ls_y, a.ls_x, y, x);
        bg = zoom_u16(a.bg + plane * a.bg_size, a.bg_cols, a.bg_y, a.bg_x, y, x);
    }
    final_store(a, p, oidx, ls, bg, zero_plane);
}

// rotation 0 / 180: an output row is an image row, so the row part of both zooms — (g * wy) for the two grid rows — is shared
// by the whole row: kFR rows per CTA keep those products in shared memory, a thread loads its column tables once and then
// only multiplies by wx and adds, in scipy's order (((g00 wy0) wx0 + (g01 wy0) wx1) + (g10 wy1) wx0) + (g11 wy1) wx1.
constexpr int kFR = 8, kFGridMax = 64;
__global__ void __launch_bounds__(256) k_lightsheet_final_rows(const FinalArgs a)
{
    __shared__ double sA[2][kFR][kFGridMax], sB[2][kFR][kFGridMax];
    __shared__ int s_y[kFR];
    __shared__ unsigned char s_zero[2][kFR];
    const size_t plane = blockIdx.z;
    const int R = a.rows, C = a.cols;
    const int i0 = blockIdx.y * kFR;
    const bool zero_plane = a.uniform_mm && a.uniform_mm[2 * plane] == ~a.uniform_mm[2 * plane + 1];
    if (threadIdx.x < kFR) {
        const int i = min(i0 + (int)threadIdx.x, a.out_rows - 1);
        int y = a.rot == 2 ? R - 1 - i : i;
        if (a.flip) y = R - 1 - y;
        s_y[threadIdx.x] = y;
        s_zero[0][threadIdx.x] = a.ls_y.zero[y];
        s_zero[1][threadIdx.x] = a.bg_y.zero[y];
    }
    __syncthreads();
    for (int gsel = 0; gsel < 2; ++gsel) {
        const ZoomAxis &zy = gsel ? a.bg_y : a.ls_y;
        const int gcols = gsel ? a.bg_cols : a.ls_cols;
        const unsigned short *g = gsel ? a.bg + plane * a.bg_size : a.ls + plane * a.ls_size;
        for (int idx = threadIdx.x; idx < kFR * gcols; idx += 256) {
            const int r = idx / gcols, c = idx - r * gcols;
            const int y = s_y[r];
            sA[gsel][r][c] = (double)g[(size_t)zy.i0[y] * gcols + c] * zy.w0[y];
            sB[gsel][r][c] = (double)g[(size_t)zy.i1[y] * gcols + c] * zy.w1[y];
        }
    }
    __syncthreads();
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= a.out_cols) return;
    const int x = a.rot == 2 ? C - 1 - j : j;
    int x0[2], x1[2];
    double w0[2], w1[2];
    bool zx[2];
#pragma unroll
    for (int gsel = 0; gsel < 2; ++gsel) {
        const ZoomAxis &z = gsel ? a.bg_x : a.ls_x;
        zx[gsel] = z.zero[x];
        x0[gsel] = z.i0[x]; x1[gsel] = z.i1[x];
        w0[gsel] = z.w0[x]; w1[gsel] = z.w1[x];
    }
#pragma unroll
    for (int r = 0; r < kFR; ++r) {
        const int i = i0 + r;
        if (i >= a.out_rows) break;
        unsigned val[2] = {0u, 0u};
        if (!zero_plane) {
#pragma unroll
            for (int gsel = 0; gsel < 2; ++gsel) {
                if (zx[gsel] || s_zero[gsel][r]) continue;
                double t = sA[gsel][r][x0[gsel]] * w0[gsel];
                t = t + sA[gsel][r][x1[gsel]] * w1[gsel];
                t = t + sB[gsel][r][x0[gsel]] * w0[gsel];
                t = t + sB[gsel][r][x1[gsel]] * w1[gsel];
                t = t > 0.0 ? t + 0.5 : 0.0;
                if (t > 65535.0) t = 65535.0;
                val[gsel] = (unsigned)t;
            }
        }
        const size_t p = plane * (size_t)R * C + (size_t)s_y[r] * C + x;
        const size_t oidx = plane * (size_t)a.out_rows * a.out_cols + (size_t)i * a.out_cols + j;
        final_store(a, p, oidx, val[0], val[1], zero_plane);
    }
}

AxisGeom axis_geom(int size, int selem, int spacing, int step)
{
    AxisGeom g;
    g.size = size;
    g.spacing = spacing;
    g.n = size / spacing;
    g.left = g.n > 0 ? (size - (g.n - 1) * spacing) / 2 : 0;
    g.hl = selem / 2;
    g.hr = selem - g.hl;
    g.step = step > 0 ? step : 1;
    return g;
}

}  // namespace

struct B2sLightsheet {
    int rows, cols, dtype;
    AxisGeom ls_y, ls_x, bg_y, bg_x;
    double qfrac, weight;
    int q_is_100, weight_is_int;
    // per-plan device tables
    int *i0[4], *i1[4];
    double *w0[4], *w1[4];
    unsigned char *zero[4];   // 0: ls y, 1: ls x, 2: bg y, 3: bg x
    int cells;                // background windows are unions of sorted cells (window a multiple of the spacing)
    CellAxis cy, cx;
};

// host-side validation shared with the geometry-only entry point; returns nullptr and sets *err on unsupported input
const char *b2s_lightsheet_check(int rows, int cols, int artifact_length, int window)
{
    if (artifact_length < 1 || window < 1) return "artifact_length and background_window_size must be positive";
    if (cols / artifact_length < 1) return "lightsheet: artifact_length exceeds the image width (empty sub-grid)";
    if (rows / 25 < 1 || cols / 25 < 1) return "lightsheet: image smaller than the background grid spacing (25)";
    if (artifact_length > 32 * 32) return "lightsheet: artifact_length > 1024 is not implemented";
    const int per_axis = (window + 1) / 2;
    if ((long long)per_axis * per_axis > 512LL * 48) return "lightsheet: background_window_size > 312 is not implemented";
    return nullptr;
}

B2sLightsheet *b2s_lightsheet_create(int rows, int cols, int dtype, int artifact_length, int window, double percentile,
                                     double weight, int weight_is_float)
{
    B2sLightsheet *L = new B2sLightsheet();
    L->rows = rows; L->cols = cols; L->dtype = dtype;
    // core.py:1333-1348: lightsheet selem (1, artifact_length), spacing = selem; background selem (w, w), spacing 25, step 2
    L->ls_y = axis_geom(rows, 1, 1, 1);
    L->ls_x = axis_geom(cols, artifact_length, artifact_length, 1);
    L->bg_y = axis_geom(rows, window, 25, 2);
    L->bg_x = axis_geom(cols, window, 25, 2);
    const double q = 100.0 * percentile;          // lightsheet_correct.py: percentile * 100
    L->qfrac = q / 100.0;
    L->q_is_100 = q == 100.0;
    L->weight = weight;
    // fast path condition: isinstance(w, float) and integer image and grids (lightsheet_correct.py:89-93); an int weight
    // takes the generic branch whose result is the same wrap-free minimum, evaluated in float64 below
    L->weight_is_int = weight_is_float && dtype != B2S_F32;
    L->cells = dtype != B2S_F32 && window % 25 == 0 && (window / 25) * (window / 25) <= 4 * 32 && getenv("B2S_LS_BRUTE") == nullptr;
    L->cy = cell_axis(L->bg_y);
    L->cx = cell_axis(L->bg_x);
    const int n_in[4] = {L->ls_y.n, L->ls_x.n, L->bg_y.n, L->bg_x.n};
    const int n_out[4] = {rows, cols, rows, cols};
    for (int k = 0; k < 4; ++k) {
        cudaMalloc(&L->i0[k], sizeof(int) * n_out[k]);
        cudaMalloc(&L->i1[k], sizeof(int) * n_out[k]);
        cudaMalloc(&L->w0[k], sizeof(double) * n_out[k]);
        cudaMalloc(&L->w1[k], sizeof(double) * n_out[k]);
        cudaMalloc(&L->zero[k], n_out[k]);
        k_zoom_axis<<<(n_out[k] + 127) / 128, 128>>>(n_in[k], n_out[k], L->i0[k], L->i1[k], L->w0[k], L->w1[k], L->zero[k]);
    }
    cudaDeviceSynchronize();
    return L;
}

void b2s_lightsheet_destroy(B2sLightsheet *L)
{
    if (!L) return;
    for (int k = 0; k < 4; ++k) { cudaFree(L->i0[k]); cudaFree(L->i1[k]); cudaFree(L->w0[k]); cudaFree(L->w1[k]); cudaFree(L->zero[k]); }
    delete L;
}

size_t b2s_lightsheet_grid_elems(const B2sLightsheet *L, int which)
{
    if (which == 2) return L->cells ? (size_t)4 * L->cy.nc * L->cx.nc * kCellCap : 0;   // sorted cell lists (uint16)
    return which == 0 ? (size_t)L->ls_y.n * L->ls_x.n : (size_t)L->bg_y.n * L->bg_x.n;
}

// mid: post-dark image (L->dtype); ls_grid / bg_grid: per-plane uint16 grids (workspace); launches 3 kernels
void b2s_launch_lightsheet(const B2sLightsheet *L, const void *mid, unsigned short *ls_grid, unsigned short *bg_grid,
                           unsigned short *cell_lists, const B2sEpilogueArgs &e, int n_planes, cudaStream_t s)
{
    GridArgs g;
    g.img = mid; g.dtype = L->dtype; g.rows = L->rows; g.cols = L->cols;
    g.qfrac = L->qfrac; g.q_is_100 = L->q_is_100;
    // lightsheet: one warp per window, up to 32 keys per lane
    g.gy = L->ls_y; g.gx = L->ls_x; g.grid = ls_grid;
    {
        const int n_win = g.gy.n * g.gx.n;
        const int per_lane = (L->ls_x.hl + L->ls_x.hr + 31) / 32;
        dim3 grid((n_win + 7) / 8, n_planes);
        if (per_lane <= 5) k_window_percentile<32, 5><<<grid, 256, 0, s>>>(g);
        else if (per_lane <= 12) k_window_percentile<32, 12><<<grid, 256, 0, s>>>(g);
        else k_window_percentile<32, 32><<<grid, 256, 0, s>>>(g);
    }
    // background: one CTA per window
    g.gy = L->bg_y; g.gx = L->bg_x; g.grid = bg_grid;
    {
        const int n_win = g.gy.n * g.gx.n;
        const int side = (L->bg_x.hl + L->bg_x.hr + 1) / 2;
        dim3 grid(n_win, n_planes);
        if (L->cells && cell_lists) {
            CellArgs ca;
            ca.img = mid; ca.dtype = L->dtype; ca.rows = L->rows; ca.cols = L->cols;
            ca.cy = L->cy; ca.cx = L->cx; ca.lists = cell_lists;
            k_sort_cells<<<dim3((4 * L->cy.nc * L->cx.nc + 7) / 8, n_planes), 256, 0, s>>>(ca);
            CellQueryArgs q;
            q.cy = L->cy; q.cx = L->cx; q.gy = L->bg_y; q.gx = L->bg_x; q.lists = cell_lists;
            q.qfrac = L->qfrac; q.q_is_100 = L->q_is_100; q.grid = bg_grid;
            const int cells_per_window = ((L->bg_y.hl + L->bg_y.hr) / 25) * ((L->bg_x.hl + L->bg_x.hr) / 25);
            const dim3 qgrid((n_win + 7) / 8, n_planes);
            if (cells_per_window <= 64) k_window_percentile_cells<2><<<qgrid, 256, 0, s>>>(q);
            else k_window_percentile_cells<4><<<qgrid, 256, 0, s>>>(q);
        } else if (L->dtype != B2S_F32 && side * side <= 256 * 40) k_window_percentile_u16<256, 20><<<grid, 256, 0, s>>>(g);
        else if (L->dtype != B2S_F32) k_window_percentile_u16<512, 24><<<grid, 512, 0, s>>>(g);
        else if (side * side <= 256 * 40) k_window_percentile<256, 40><<<grid, 256, 0, s>>>(g);
        else k_window_percentile<512, 48><<<grid, 512, 0, s>>>(g);
    }
    FinalArgs f;
    f.img = mid; f.dtype = L->dtype; f.rows = L->rows; f.cols = L->cols;
    f.ls = ls_grid; f.bg = bg_grid;
    f.ls_cols = L->ls_x.n; f.bg_cols = L->bg_x.n;
    f.ls_size = L->ls_y.n * L->ls_x.n; f.bg_size = L->bg_y.n * L->bg_x.n;
    const ZoomAxis za[4] = {{L->i0[0], L->i1[0], L->w0[0], L->w1[0], L->zero[0]}, {L->i0[1], L->i1[1], L->w0[1], L->w1[1], L->zero[1]},
                            {L->i0[2], L->i1[2], L->w0[2], L->w1[2], L->zero[2]}, {L->i0[3], L->i1[3], L->w0[3], L->w1[3], L->zero[3]}};
    f.ls_y = za[0]; f.ls_x = za[1]; f.bg_y = za[2]; f.bg_x = za[3];
    f.weight = L->weight; f.weight_is_int = L->weight_is_int;
    f.final_mode = e.final_mode; f.shift = e.shift; f.out_dtype = e.out_dtype; f.flip = e.flip; f.rot = e.rot;
    f.uniform_mm = e.uniform_mm; f.out = e.out; f.out_rows = e.out_rows; f.out_cols = e.out_cols;
    if ((e.rot == 0 || e.rot == 2) && f.ls_cols <= kFGridMax && f.bg_cols <= kFGridMax)
        k_lightsheet_final_rows<<<dim3((e.out_cols + 255) / 256, (e.out_rows + kFR - 1) / kFR, n_planes), 256, 0, s>>>(f);
    else
        k_lightsheet_final<<<dim3((e.out_cols + 255) / 256, e.out_rows, n_planes), 256, 0, s>>>(f);
}
