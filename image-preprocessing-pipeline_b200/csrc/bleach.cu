// correct_bleaching (reference pystripe/core.py:501-559, non-max method) and butter_lowpass_filter (core.py:493-499) on
// the cropped log-domain image, between the reconstruction and expm1 (core.py:1131-1148).
//
//   img_filter = img.copy(); img_filter[img_filter == 0] = clip_med; clip(img_filter, clip_min, clip_max)
//   img_filter = sosfiltfilt(butter(1, frequency, output='sos'), img_filter).astype(float32)     # along each row
//   img = img / img_filter * max(img_filter)                                                      # float32
//
// scipy.signal.sosfiltfilt, restated operation for operation (checked bit-for-bit against scipy on the CPU by
// tests/test_oracle.py::test_sosfiltfilt_restatement and on the GPU against the reference run verbatim):
//   * odd extension by edge = 6 samples per side, evaluated in the image's float32 (2*x[0] - x[6-j], 2*x[n-1] - x[n-2-k]);
//   * one second-order section [b0, b1, 0, 1, a1, 0] in float64, direct form II transposed as _sosfilt.pyx runs it:
//       y = b0*x + z;  z = (b1*x - a1*y) [+ 0]        (separate multiplies and adds: file compiled with -fmad=false)
//     started at z = zi * x_ext[0]; the output is reversed and filtered again from z = zi * y[last]; the edges are
//     dropped and the result is cast to float32.
// The recurrence is sequential along a row, rows are independent: one thread per row, 32 rows per CTA run by one consumer
// warp; three producer warps move the row segments through double-buffered shared-memory tiles so that global
// loads/stores stay coalesced and overlap the recurrence.  The forward output (float64) goes through a per-slot scratch
// buffer that the same CTA reads back.
#include "b2s_internal.h"

namespace {

constexpr int kRows = 32;      // rows (= threads) per CTA: one warp, so that 2048 rows x 8 planes already give 512 CTAs
constexpr int kChunk = 32;     // samples per tile
constexpr int kEdge = 6;       // sosfiltfilt: ntaps = 2*1 + 1 - min(#(b2 == 0), #(a2 == 0)) = 2; edge = 3 * ntaps

__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// the value the low-pass filter sees at column c of a row (float32): zero -> clip_med, then numpy.clip
__device__ __forceinline__ float filt_input(const float *row, int c, const B2sBleachArgs &a)
{
    float v = row[c];
    if (v == 0.0f) v = (float)a.clip_med;
    double d = (double)v;
    d = d < a.clip_min ? a.clip_min : d;      // numpy.clip = minimum(maximum(x, lo), hi)
    d = d > a.clip_max ? a.clip_max : d;
    return (float)d;
}

// sample j of the odd-extended row (float32 arithmetic, scipy.signal._arraytools.odd_ext)
__device__ __forceinline__ float ext_sample(const float *row, int j, int n, const B2sBleachArgs &a)
{
    if (j < kEdge) return __fsub_rn(2.0f * filt_input(row, 0, a), filt_input(row, kEdge - j, a));
    if (j < kEdge + n) return filt_input(row, j - kEdge, a);
    return __fsub_rn(2.0f * filt_input(row, n - 1, a), filt_input(row, n - 2 - (j - kEdge - n), a));
}

// One CTA = 32 rows: warp 0 runs the recurrences (lane = row), warps 1-3 move the tiles.  Tiles are double-buffered in
// shared memory, one barrier per tile: while the consumer warp works through tile k, the producer warps fetch tile k+1
// from global memory (clip / odd extension applied on the way) and write the consumer's tile k-1 back, so the time per
// tile is max(recurrence, memory) instead of their sum.
constexpr int kProducers = 96;
constexpr int kTileElems = kRows * kChunk;
constexpr int kPerProducer = (kTileElems + kProducers - 1) / kProducers;   // 11

__global__ void __launch_bounds__(kRows + kProducers) k_bleach_lowpass(const B2sBleachArgs a_in)
{
    B2sBleachArgs a = a_in;
    if (a.clip_pp) {   // per-plane clip levels (multi-Otsu)
        const double *cl = a.clip_pp + 3 * (size_t)blockIdx.y;
        a.clip_min = cl[0]; a.clip_med = cl[1]; a.clip_max = cl[2];
    }
    static_assert(kRows == 32 && kChunk == 32, "warp 0: lane = row of the tile");
    __shared__ float s_f[2][kRows][kChunk + 1];
    __shared__ double s_d[2][kRows][kChunk + 1];
    const int tid = threadIdx.x;
    const bool consumer = tid < kRows;
    const int ptid = tid - kRows;
    const size_t plane = blockIdx.y;
    const int r0 = blockIdx.x * kRows;
    const int n = a.cols, N = n + 2 * kEdge;
    const int nch = (N + kChunk - 1) / kChunk;
    const float *img = a.img.ptr + plane * a.img.plane_stride + (size_t)a.base_pad * a.img.pitch + a.base_pad;
    double *scr = a.scratch + plane * a.scratch_plane_stride;
    float *filt = a.filt + plane * (size_t)a.rows * n;
    const bool live = consumer && r0 + tid < a.rows;
    const double b0 = a.b0, b1 = a.b1, a1 = a.a1;
    double z = 0.0, last = 0.0;

    auto load_ext = [&](int t, int buf) {            // producers: tile t of the extended rows -> s_f[buf]
        float v[kPerProducer];
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i, rr = e / kChunk, jj = e % kChunk, r = r0 + rr, j = t * kChunk + jj;
            v[i] = (e < kTileElems && r < a.rows && j < N) ? ext_sample(img + (size_t)r * a.img.pitch, j, n, a) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i;
            if (e < kTileElems) s_f[buf][e / kChunk][e % kChunk] = v[i];
        }
    };
    auto store_scr = [&](int t, int buf) {           // producers: s_d[buf] -> scratch tile t
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i, rr = e / kChunk, jj = e % kChunk, r = r0 + rr, j = t * kChunk + jj;
            if (e < kTileElems && r < a.rows && j < N) scr[(size_t)r * N + j] = s_d[buf][rr][jj];
        }
    };
    auto load_scr = [&](int t, int buf) {            // producers: scratch tile t -> s_d[buf]
        double v[kPerProducer];
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i, rr = e / kChunk, jj = e % kChunk, r = r0 + rr, j = t * kChunk + jj;
            v[i] = (e < kTileElems && r < a.rows && j < N) ? scr[(size_t)r * N + j] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i;
            if (e < kTileElems) s_d[buf][e / kChunk][e % kChunk] = v[i];
        }
    };
    auto store_filt = [&](int t, int buf) {          // producers: s_f[buf] -> img_filter tile t (edges dropped)
#pragma unroll
        for (int i = 0; i < kPerProducer; ++i) {
            const int e = ptid + kProducers * i, rr = e / kChunk, jj = e % kChunk, r = r0 + rr, j = t * kChunk + jj;
            if (e < kTileElems && r < a.rows && j >= kEdge && j < kEdge + n) filt[(size_t)r * n + (j - kEdge)] = s_f[buf][rr][jj];
        }
    };

    // forward over the extended row
    if (!consumer) load_ext(0, 0);
    __syncthreads();
    for (int k = 0; k < nch; ++k) {
        if (!consumer) {
            if (k + 1 < nch) load_ext(k + 1, (k + 1) & 1);
            if (k >= 1) store_scr(k - 1, (k - 1) & 1);
        } else if (live) {
            const int b = k & 1;
            if (k == 0) z = a.zi * (double)s_f[0][tid][0];
            const int m = min(kChunk, N - k * kChunk);
#pragma unroll 8
            for (int jj = 0; jj < m; ++jj) {
                const double x = (double)s_f[b][tid][jj];
                const double y = __dadd_rn(__dmul_rn(b0, x), z);
                z = __dsub_rn(__dmul_rn(b1, x), __dmul_rn(a1, y));
                s_d[b][tid][jj] = y;
                last = y;
            }
        }
        __syncthreads();
    }
    if (!consumer) store_scr(nch - 1, (nch - 1) & 1);
    __syncthreads();     // block-wide: the scratch writes above are visible to this CTA's reads below

    // backward: the reversed forward output through the same section, started from zi * y[last]
    z = a.zi * last;
    float mx = -INFINITY;
    if (!consumer) load_scr(nch - 1, 0);
    __syncthreads();
    for (int q = 0; q < nch; ++q) {
        const int t = nch - 1 - q;
        if (!consumer) {
            if (q + 1 < nch) load_scr(t - 1, (q + 1) & 1);
            if (q >= 1) store_filt(t + 1, (q - 1) & 1);
        } else if (live) {
            const int b = q & 1;
            const int m = min(kChunk, N - t * kChunk);
#pragma unroll 8
            for (int jj = m - 1; jj >= 0; --jj) {
                const double x = s_d[b][tid][jj];
                const double y = __dadd_rn(__dmul_rn(b0, x), z);
                z = __dsub_rn(__dmul_rn(b1, x), __dmul_rn(a1, y));
                const float f = (float)y;
                s_f[b][tid][jj] = f;
                const int j = t * kChunk + jj;
                if (j >= kEdge && j < kEdge + n) mx = fmaxf(mx, f);
            }
        }
        __syncthreads();
    }
    if (!consumer) store_filt(0, (nch - 1) & 1);
    if (consumer) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        if (tid == 0) atomicMax(a.maxkey + plane, f2key(mx));
    }
}

// img = img / img_filter * max(img_filter), float32, in place on the cropped window of the padded image
__global__ void __launch_bounds__(256) k_bleach_apply(B2sBleachArgs a)
{
    const size_t plane = blockIdx.z;
    const int r = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= a.cols) return;
    float *p = a.img.ptr + plane * a.img.plane_stride + (size_t)(a.base_pad + r) * a.img.pitch + a.base_pad + c;
    const float f = a.filt[plane * (size_t)a.rows * a.cols + (size_t)r * a.cols + c];
    const float mx = key2f(a.maxkey[plane]);
    *p = __fmul_rn(__fdiv_rn(*p, f), mx);
}

// ---- max method (core.py:533-545): the filter is the outer product of the low-passed row maxima and column maxima
__global__ void __launch_bounds__(256) k_bleach_row_max(B2sBleachArgs a, float *vy)
{
    __shared__ float s_red[8];
    const size_t plane = blockIdx.y;
    const int r = blockIdx.x;
    const float *row = a.img.ptr + plane * a.img.plane_stride + (size_t)(a.base_pad + r) * a.img.pitch + a.base_pad;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < a.cols; c += 256) m = fmaxf(m, row[c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
        vy[plane * a.rows + r] = m;
    }
}

__global__ void __launch_bounds__(256) k_bleach_col_max(B2sBleachArgs a, float *vx)
{
    const size_t plane = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= a.cols) return;
    const float *p = a.img.ptr + plane * a.img.plane_stride + (size_t)a.base_pad * a.img.pitch + a.base_pad + c;
    float m = -INFINITY;
    for (int r = 0; r < a.rows; ++r) m = fmaxf(m, p[(size_t)r * a.img.pitch]);
    vx[plane * a.cols + c] = m;
}

// max over the outer product (float32 products, as numpy.dot of an (H, 1) by a (1, W) matrix forms them)
__global__ void __launch_bounds__(256) k_bleach_outer_max(B2sBleachArgs a, const float *fy, const float *fx)
{
    __shared__ float s_red[8];
    const size_t plane = blockIdx.y;
    const float y = fy[plane * a.rows + blockIdx.x];
    float m = -INFINITY;
    for (int c = threadIdx.x; c < a.cols; c += 256) m = fmaxf(m, __fmul_rn(y, fx[plane * a.cols + c]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
        atomicMax(a.maxkey + plane, f2key(m));
    }
}

__global__ void __launch_bounds__(256) k_bleach_apply_outer(B2sBleachArgs a, const float *fy, const float *fx)
{
    const size_t plane = blockIdx.z;
    const int r = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= a.cols) return;
    float *p = a.img.ptr + plane * a.img.plane_stride + (size_t)(a.base_pad + r) * a.img.pitch + a.base_pad + c;
    const float f = __fmul_rn(fy[plane * a.rows + r], fx[plane * a.cols + c]);
    *p = __fmul_rn(__fdiv_rn(*p, f), key2f(a.maxkey[plane]));
}

}  // namespace

// max method: a.filt holds, per plane, [row maxima (rows) | column maxima (cols) | filtered rows | filtered cols];
// a.scratch needs (max(rows, cols) + 12) doubles per plane
void b2s_launch_bleach_max_method(const B2sBleachArgs &a, int n_planes, cudaStream_t s)
{
    const size_t per = 2 * ((size_t)a.rows + a.cols);
    float *vy = a.filt, *vx = vy + (size_t)n_planes * a.rows;
    float *fy = vx + (size_t)n_planes * a.cols, *fx = fy + (size_t)n_planes * a.rows;
    (void)per;
    k_bleach_row_max<<<dim3(a.rows, n_planes), 256, 0, s>>>(a, vy);
    k_bleach_col_max<<<dim3((a.cols + 255) / 256, n_planes), 256, 0, s>>>(a, vx);
    // the 1-D low-pass of each vector = the 2-D kernel on a one-row image per plane
    for (int which = 0; which < 2; ++which) {
        B2sBleachArgs v = a;
        const int len = which == 0 ? a.rows : a.cols;
        v.img.ptr = which == 0 ? vy : vx;
        v.img.plane_stride = len; v.img.pitch = len; v.img.rows = 1; v.img.cols = len;
        v.base_pad = 0; v.rows = 1; v.cols = len;
        v.scratch_plane_stride = (size_t)len + 12;
        v.filt = which == 0 ? fy : fx;
        cudaMemsetAsync(a.maxkey, 0, sizeof(unsigned) * n_planes, s);
        k_bleach_lowpass<<<dim3(1, n_planes), kRows + kProducers, 0, s>>>(v);
    }
    cudaMemsetAsync(a.maxkey, 0, sizeof(unsigned) * n_planes, s);
    k_bleach_outer_max<<<dim3(a.rows, n_planes), 256, 0, s>>>(a, fy, fx);
    k_bleach_apply_outer<<<dim3((a.cols + 255) / 256, a.rows, n_planes), 256, 0, s>>>(a, fy, fx);
}

void b2s_launch_bleach(const B2sBleachArgs &a, int n_planes, cudaStream_t s)
{
    cudaMemsetAsync(a.maxkey, 0, sizeof(unsigned) * n_planes, s);
    k_bleach_lowpass<<<dim3((a.rows + kRows - 1) / kRows, n_planes), kRows + kProducers, 0, s>>>(a);
    k_bleach_apply<<<dim3((a.cols + 255) / 256, a.rows, n_planes), 256, 0, s>>>(a);
}
