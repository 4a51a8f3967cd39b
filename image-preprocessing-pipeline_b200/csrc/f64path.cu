// Double-precision destripe for integer pixels WITHOUT log normalisation.
//
// filter_streaks(img, log1p_normalization_needed=False) on a uint8 / uint16 image (pystripe/core.py:1063, 1081-1158) never
// leaves the integer dtype before the wavelet transform, so PyWavelets promotes it to float64 and everything up to the
// final rint / clip runs in double precision: np.pad on the integers, pywt.wavedec2 / waverec2 (float64), scipy.fftpack
// rfft / irfft (float64) with the float32 notch broadcast into it.  A float64 pipeline is accurate to ~1e-9 of a grey level
// here, and every pass ends with `waverec2(...).astype(d_type)` (filter_subband, core.py:939: a C cast back to the integer
// dtype — truncation toward zero, wrap-around outside the range), so this path does NOT mirror the reference's operation
// order: any double-precision evaluation gives the same integers except when a value lies within ~1e-9 of an integer
// (probability ~1e-8 per pixel; the reference's own result is just as fragile there).  It is a completeness path (no caller of the pipeline passes log1p False), written for
// clarity: one thread per output, separable passes through global memory, and the notch  irfft(rfft(x) * g)  applied as the
// dense real n x n matrix it is (the response to every unit vector, evaluated by scipy on the host: b2s_plan_set_notch_matrix).
#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

__device__ __forceinline__ int imod(int i, int p)
{
    int t = i % p;
    return t < 0 ? t + p : t;
}
__device__ __forceinline__ int pad_index(int i, int n, int mode)   // numpy.pad source index, -1 = constant fill
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case B2S_PAD_REFLECT: { if (n == 1) return 0; const int p = 2 * (n - 1), t = imod(i, p); return t < n ? t : p - t; }
    case B2S_PAD_SYMMETRIC: { const int p = 2 * n, t = imod(i, p); return t < n ? t : p - 1 - t; }
    case B2S_PAD_WRAP: return imod(i, n);
    case B2S_PAD_EDGE: return i < 0 ? 0 : n - 1;
    default: return -1;
    }
}
__device__ __forceinline__ int sym_ext(int i, int n)   // pywt MODE_SYMMETRIC (half-sample), any i
{
    if (i >= 0 && i < n) return i;
    const int p = 2 * n, t = imod(i, p);
    return t < n ? t : p - 1 - t;
}

struct Taps64 { int F; double dec_lo[B2S_MAX_TAPS], dec_hi[B2S_MAX_TAPS], rec_lo[B2S_MAX_TAPS], rec_hi[B2S_MAX_TAPS]; };

__global__ void k_f64_pad(const void *in, int dtype, int rows, int cols, int mode, int base, double fill, double *out, int PH,
                          int PW, size_t out_stride, const unsigned char *mask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= PW) return;
    const size_t plane = blockIdx.z;
    const int sy = pad_index(y - base, rows, mode), sx = pad_index(x - base, cols, mode);
    double v = fill;
    if (sy >= 0 && sx >= 0) {
        const size_t idx = plane * (size_t)rows * cols + (size_t)sy * cols + sx;
        v = dtype == B2S_U16 ? (double)reinterpret_cast<const unsigned short *>(in)[idx]
                             : (double)reinterpret_cast<const unsigned char *>(in)[idx];
        if (mask && !mask[idx]) v = 0.0;       // img *= mask (core.py:1080) on the integer image
    }
    out[plane * out_stride + (size_t)y * PW + x] = v;
}

// analysis along axis -2: in (ny x nx) -> lo, hi (my x nx)
__global__ void k_f64_fwd_y(const __grid_constant__ Taps64 t, const double *in, int ny, int nx, size_t in_stride, double *lo,
                            double *hi, int my, size_t out_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
    if (x >= nx) return;
    const double *src = in + blockIdx.z * in_stride;
    double a = 0.0, b = 0.0;
    for (int j = 0; j < t.F; ++j) {
        const double v = src[(size_t)sym_ext(2 * o + 1 - j, ny) * nx + x];
        a += t.dec_lo[j] * v;
        b += t.dec_hi[j] * v;
    }
    lo[blockIdx.z * out_stride + (size_t)o * nx + x] = a;
    hi[blockIdx.z * out_stride + (size_t)o * nx + x] = b;
}
// analysis along axis -1: in (my x nx) -> lo, hi (my x mx)
__global__ void k_f64_fwd_x(const __grid_constant__ Taps64 t, const double *in, int my, int nx, size_t in_stride, double *lo,
                            double *hi, int mx, size_t out_stride)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (o >= mx) return;
    const double *src = in + blockIdx.z * in_stride + (size_t)y * nx;
    double a = 0.0, b = 0.0;
    for (int j = 0; j < t.F; ++j) {
        const double v = src[sym_ext(2 * o + 1 - j, nx)];
        a += t.dec_lo[j] * v;
        b += t.dec_hi[j] * v;
    }
    lo[blockIdx.z * out_stride + (size_t)y * mx + o] = a;
    hi[blockIdx.z * out_stride + (size_t)y * mx + o] = b;
}
// synthesis along axis -1: (a, d) (my x mx) -> out (my x ox), ox <= 2 mx - F + 2
__global__ void k_f64_inv_x(const __grid_constant__ Taps64 t, const double *ca, const double *cd, int my, int mx, size_t in_stride,
                            double *out, int ox, size_t out_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ox) return;
    const int H = t.F / 2, p = x >> 1, e = x & 1;
    const double *pa = ca + blockIdx.z * in_stride + (size_t)y * mx, *pd = cd + blockIdx.z * in_stride + (size_t)y * mx;
    double s = 0.0;
    for (int j = 0; j < H; ++j) {
        const int c = p + H - 1 - j;
        if (c < mx) s += t.rec_lo[2 * j + e] * pa[c] + t.rec_hi[2 * j + e] * pd[c];
    }
    out[blockIdx.z * out_stride + (size_t)y * ox + x] = s;
}
// synthesis along axis -2: (a, d) (my x nx) -> out (oy x nx), oy <= 2 my - F + 2
__global__ void k_f64_inv_y(const __grid_constant__ Taps64 t, const double *ca, const double *cd, int my, int nx, size_t in_stride,
                            double *out, int oy, size_t out_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const int H = t.F / 2, p = y >> 1, e = y & 1;
    const double *pa = ca + blockIdx.z * in_stride + x, *pd = cd + blockIdx.z * in_stride + x;
    double s = 0.0;
    for (int j = 0; j < H; ++j) {
        const int c = p + H - 1 - j;
        if (c < my) s += t.rec_lo[2 * j + e] * pa[(size_t)c * nx] + t.rec_hi[2 * j + e] * pd[(size_t)c * nx];
    }
    out[blockIdx.z * out_stride + (size_t)y * nx + x] = s;
}

// out = in * R (rows transformed: out[r][c] = sum_k in[r][k] R[k][c]) or out = R^T * in (columns: out[r][c] = sum_k R[k][r] in[k][c]);
// 32 x 32 output tile per CTA, shared-memory tiles of 32
__global__ void __launch_bounds__(256) k_f64_notch(const double *in, const double *R, double *out, int rows, int cols, int along_cols,
                                                   size_t stride)
{
    __shared__ double sa[32][33], sb[32][33];
    const double *src = in + blockIdx.z * stride;
    double *dst = out + blockIdx.z * stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads, 4 outputs each
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int n = along_cols ? rows : cols;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k0 = 0; k0 < n; k0 += 32) {
        for (int q = 0; q < 4; ++q) {
            const int i = ty + 8 * q;
            if (!along_cols) {
                // sa[i][tx] = in[r0 + i][k0 + tx];  sb[i][tx] = R[k0 + i][c0 + tx]
                sa[i][tx] = (r0 + i < rows && k0 + tx < n) ? src[(size_t)(r0 + i) * cols + k0 + tx] : 0.0;
                sb[i][tx] = (k0 + i < n && c0 + tx < cols) ? R[(size_t)(k0 + i) * n + c0 + tx] : 0.0;
            } else {
                // sa[i][tx] = R[k0 + tx][r0 + i]  (transposed use);  sb[i][tx] = in[k0 + i][c0 + tx]
                sa[i][tx] = (r0 + i < rows && k0 + tx < n) ? R[(size_t)(k0 + tx) * n + r0 + i] : 0.0;
                sb[i][tx] = (k0 + i < n && c0 + tx < cols) ? src[(size_t)(k0 + i) * cols + c0 + tx] : 0.0;
            }
        }
        __syncthreads();
        for (int k = 0; k < 32; ++k) {
            const double b = sb[k][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += sa[ty + 8 * q][k] * b;
        }
        __syncthreads();
    }
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty + 8 * q, c = c0 + tx;
        if (r < rows && c < cols) dst[(size_t)r * cols + c] = acc[q];
    }
}

// end of a pass: `.astype(d_type)` of the float64 reconstruction (core.py:939) — numpy's C cast: truncation toward zero,
// then wrap-around modulo 2^bits (what the x86 conversion through a wider integer gives for out-of-range values)
__global__ void k_f64_astype_int(double *img, size_t n_per_plane, size_t stride, long long mask)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_plane) return;
    double *p = img + blockIdx.y * stride + i;
    const long long t = (long long)(*p);        // cvttsd2si: toward zero
    *p = (double)(t & mask);
}

// reconstruction (double, PH x PW) -> the float32 working image the common epilogue reads: rint (half to even), clip to the
// integer dtype (core.py:1153-1158); the values are integers <= 65535, exact in float32
__global__ void k_f64_to_f32(const double *in, int PH, int PW, size_t in_stride, float *out, int pitch, size_t out_stride, double hi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= PW) return;
    double v = rint(in[blockIdx.z * in_stride + (size_t)y * PW + x]);
    v = v < 0.0 ? 0.0 : (v > hi ? hi : v);
    out[blockIdx.z * out_stride + (size_t)y * pitch + x] = (float)v;
}

}  // namespace

size_t b2s_f64_workspace_doubles(int PH, int PW, int levels, const int *my, const int *mx)
{
    // padded image + four sub-bands per level + two half-transformed intermediates of level 1 + one notch output buffer
    size_t n = (size_t)PH * PW;
    for (int l = 1; l <= levels; ++l) n += 4 * (size_t)my[l] * mx[l];
    n += 2 * (size_t)my[1] * PW;
    n += (size_t)my[1] * mx[1];
    return n;
}

void b2s_launch_f64_destripe(const B2sF64Args &a, int n_planes, cudaStream_t s)
{
    Taps64 t;
    t.F = a.F;
    for (int k = 0; k < a.F; ++k) { t.dec_lo[k] = a.dec_lo[k]; t.dec_hi[k] = a.dec_hi[k]; t.rec_lo[k] = a.rec_lo[k]; t.rec_hi[k] = a.rec_hi[k]; }
    const size_t st = a.plane_doubles;   // every per-plane buffer is laid out inside one block of `st` doubles per plane
    // offsets inside a plane's block
    size_t off = 0;
    const size_t o_pad = off; off += (size_t)a.PH * a.PW;
    size_t o_sub[B2S_MAX_LEVELS + 1][4];
    for (int l = 1; l <= a.levels; ++l)
        for (int k = 0; k < 4; ++k) { o_sub[l][k] = off; off += (size_t)a.my[l] * a.mx[l]; }
    const size_t o_lo = off; off += (size_t)a.my[1] * a.PW;
    const size_t o_hi = off; off += (size_t)a.my[1] * a.PW;
    const size_t o_tmp = off;
    double *W = a.work;
    const dim3 blk(256);
    k_f64_pad<<<dim3((a.PW + 255) / 256, a.PH, n_planes), blk, 0, s>>>(a.in, a.in_dtype, a.rows, a.cols, a.pad_mode, a.base_pad,
                                                                      a.pad_value, W + o_pad, a.PH, a.PW, st, a.mask);
    for (int pass = 0; pass < a.n_passes; ++pass) {
        for (int l = 1; l <= a.levels; ++l) {   // sub order: 0 = cA, 1 = cH (rows hi, cols lo), 2 = cV (rows lo, cols hi), 3 = cD
            const int ny = a.my[l - 1], nx = a.mx[l - 1], my = a.my[l], mx = a.mx[l];
            const double *src = l == 1 ? W + o_pad : W + o_sub[l - 1][0];
            k_f64_fwd_y<<<dim3((nx + 255) / 256, my, n_planes), blk, 0, s>>>(t, src, ny, nx, st, W + o_lo, W + o_hi, my, st);
            k_f64_fwd_x<<<dim3((mx + 255) / 256, my, n_planes), blk, 0, s>>>(t, W + o_lo, my, nx, st, W + o_sub[l][0], W + o_sub[l][2], mx, st);
            k_f64_fwd_x<<<dim3((mx + 255) / 256, my, n_planes), blk, 0, s>>>(t, W + o_hi, my, nx, st, W + o_sub[l][1], W + o_sub[l][3], mx, st);
        }
        for (int l = 1; l <= a.levels; ++l) {
            const int my = a.my[l], mx = a.mx[l];
            const dim3 grid((mx + 31) / 32, (my + 31) / 32, n_planes);
            k_f64_notch<<<grid, 256, 0, s>>>(W + o_sub[l][1], a.notch[pass][l][0], W + o_tmp, my, mx, 0, st);
            cudaMemcpy2DAsync(W + o_sub[l][1], sizeof(double) * st, W + o_tmp, sizeof(double) * st, sizeof(double) * (size_t)my * mx,
                              n_planes, cudaMemcpyDeviceToDevice, s);
            if (a.bidirectional) {
                k_f64_notch<<<grid, 256, 0, s>>>(W + o_sub[l][2], a.notch[pass][l][1], W + o_tmp, my, mx, 1, st);
                cudaMemcpy2DAsync(W + o_sub[l][2], sizeof(double) * st, W + o_tmp, sizeof(double) * st, sizeof(double) * (size_t)my * mx,
                                  n_planes, cudaMemcpyDeviceToDevice, s);
            }
        }
        for (int l = a.levels; l >= 1; --l) {
            const int my = a.my[l], mx = a.mx[l], oy = a.my[l - 1], ox = a.mx[l - 1];
            double *dst = l == 1 ? W + o_pad : W + o_sub[l - 1][0];
            // axis -1 first: (cA, cV) -> a, (cH, cD) -> d (my x ox), then axis -2
            k_f64_inv_x<<<dim3((ox + 255) / 256, my, n_planes), blk, 0, s>>>(t, W + o_sub[l][0], W + o_sub[l][2], my, mx, st, W + o_lo, ox, st);
            k_f64_inv_x<<<dim3((ox + 255) / 256, my, n_planes), blk, 0, s>>>(t, W + o_sub[l][1], W + o_sub[l][3], my, mx, st, W + o_hi, ox, st);
            k_f64_inv_y<<<dim3((ox + 255) / 256, oy, n_planes), blk, 0, s>>>(t, W + o_lo, W + o_hi, my, ox, st, dst, oy, st);
        }
        const size_t npp = (size_t)a.PH * a.PW;
        k_f64_astype_int<<<dim3((unsigned)((npp + 255) / 256), n_planes), blk, 0, s>>>(W + o_pad, npp, st, a.in_dtype == B2S_U8 ? 0xffLL : 0xffffLL);
    }
    k_f64_to_f32<<<dim3((a.PW + 255) / 256, a.PH, n_planes), blk, 0, s>>>(W + o_pad, a.PH, a.PW, st, a.out.ptr, a.out.pitch,
                                                                         a.out.plane_stride, a.in_dtype == B2S_U8 ? 255.0 : 65535.0);
}
