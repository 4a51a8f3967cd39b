// get_img_mask (pystripe/core.py:475-489) and `img *= mask` (core.py:1079-1080): the foreground mask filter_streaks
// multiplies into the (log) image before it is padded when enable_masking is set.
//
//   mask = img > threshold                                     (core.py:479)
//   mask = cv2.morphologyEx(mask, MORPH_CLOSE, ones(c, c))      dilate then erode, same kernel and anchor   (core.py:480)
//   mask = cv2.morphologyEx(mask, MORPH_OPEN,  ones(o, o))      erode then dilate                           (core.py:481)
//   holes = ~mask with the 4-connected components that contain a corner pixel removed (four cv2.floodFill) (core.py:482-486)
//   mask |= holes                                                                                          (core.py:487)
//
// OpenCV's rectangular k x k element with the default anchor (k/2, k/2) covers the offsets [-(k/2), k-1-(k/2)] on both
// axes for dilate AND erode (no reflection for even k); outside the image dilate sees 0 and erode sees 1
// (morphologyDefaultBorderValue).  A k x k rectangle is separable and the image is binary, so each pass is a window count:
// along rows through a shared-memory prefix sum (one CTA per row), along columns as a sliding count (one thread per
// column, coalesced).  The flood fills are a reachability fixed point: sweeps along rows and columns repeated until no
// pixel changes (the host polls a flag every few rounds).  Nothing here is on the hot path: no caller of the pipeline
// enables masking — the kernels are written for clarity.
#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

typedef unsigned char u8;

// ---- threshold -------------------------------------------------------------------------------------------------------
// the comparison runs in double: float32 / integer pixels convert exactly, and the caller hands the threshold in the
// precision numpy compares in (a weak Python scalar rounded to float32 against a float32 image, etc.)
__global__ void k_mask_threshold_f32(B2sImg img, int base_pad, int rows, int cols, double thr, const double *thr_pp, u8 *mask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const size_t plane = blockIdx.z;
    const double t = thr_pp ? thr_pp[plane] : thr;
    const float v = img.ptr[plane * img.plane_stride + (size_t)(y + base_pad) * img.pitch + base_pad + x];
    mask[plane * (size_t)rows * cols + (size_t)y * cols + x] = (double)v > t ? 1 : 0;
}
template <typename T>
__global__ void k_mask_threshold_int(const T *in, size_t n_per_plane, double thr, const double *thr_pp, u8 *mask)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_plane) return;
    const size_t plane = blockIdx.y;
    const double t = thr_pp ? thr_pp[plane] : thr;
    mask[plane * n_per_plane + i] = (double)in[plane * n_per_plane + i] > t ? 1 : 0;
}

// ---- morphology ------------------------------------------------------------------------------------------------------
// window [x + lo, x + hi] clipped to the line; dilate: any one inside; erode: no zero inside (outside counts as one)
__global__ void __launch_bounds__(256) k_morph_rows(const u8 *in, u8 *out, int rows, int cols, int lo, int hi, int erode)
{
    extern __shared__ int s_pre[];      // s_pre[x] = ones in [0, x)
    __shared__ int s_part[256];
    const int y = blockIdx.x;
    const size_t off = blockIdx.y * (size_t)rows * cols + (size_t)y * cols;
    const int chunk = (cols + 255) / 256;
    const int x0 = threadIdx.x * chunk, x1 = min(cols, x0 + chunk);
    int sum = 0;
    for (int x = x0; x < x1; ++x) sum += in[off + x];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < 256; ++i) { const int t = s_part[i]; s_part[i] = run; run += t; }
    }
    __syncthreads();
    int run = s_part[threadIdx.x];
    for (int x = x0; x < x1; ++x) { s_pre[x] = run; run += in[off + x]; }
    if (x1 == cols && x0 < cols) s_pre[cols] = run;
    if (cols == 0) return;
    __syncthreads();
    for (int x = threadIdx.x; x < cols; x += 256) {
        const int a = max(0, x + lo), b = min(cols - 1, x + hi);
        const int ones = b >= a ? s_pre[b + 1] - s_pre[a] : 0;
        out[off + x] = erode ? (ones == (b >= a ? b - a + 1 : 0) ? 1 : 0) : (ones > 0 ? 1 : 0);
    }
}
__global__ void k_morph_cols(const u8 *in, u8 *out, int rows, int cols, int lo, int hi, int erode)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cols) return;
    const size_t off = blockIdx.y * (size_t)rows * cols + x;
    int ones = 0;                                  // ones in [y + lo, y + hi] clipped
    for (int r = max(0, lo); r <= min(rows - 1, hi); ++r) ones += in[off + (size_t)r * cols];
    for (int y = 0; y < rows; ++y) {
        const int a = max(0, y + lo), b = min(rows - 1, y + hi);
        const int len = b >= a ? b - a + 1 : 0;
        out[off + (size_t)y * cols] = erode ? (ones == len ? 1 : 0) : (ones > 0 ? 1 : 0);
        const int leave = y + lo, enter = y + 1 + hi;
        if (leave >= 0 && leave < rows) ones -= in[off + (size_t)leave * cols];
        if (enter >= 0 && enter < rows) ones += in[off + (size_t)enter * cols];
    }
}

// ---- the four flood fills: background reachable from a corner pixel through 4-connected background ---------------------
__global__ void k_flood_init(const u8 *mask, u8 *reach, int rows, int cols)
{
    const size_t n = (size_t)rows * cols, off = blockIdx.x * n;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) reach[off + i] = 0;
    __syncthreads();
    if (threadIdx.x == 0 && n > 0) {
        const size_t c[4] = {0, (size_t)cols - 1, (size_t)(rows - 1) * cols, n - 1};
        for (int k = 0; k < 4; ++k)
            if (!mask[off + c[k]]) reach[off + c[k]] = 1;
    }
}
// one thread per line: forward then backward sweep; stride 1 along rows, `cols` along columns
__global__ void k_flood_sweep(const u8 *mask, u8 *reach, int rows, int cols, int along_cols, int *changed)
{
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_lines = along_cols ? cols : rows, len = along_cols ? rows : cols;
    if (line >= n_lines) return;
    const size_t off = blockIdx.y * (size_t)rows * cols + (along_cols ? (size_t)line : (size_t)line * cols);
    const size_t step = along_cols ? (size_t)cols : 1;
    bool any = false;
    u8 r = 0;
    for (int i = 0; i < len; ++i) {
        const size_t p = off + i * step;
        const u8 bg = mask[p] ? 0 : 1, cur = reach[p];
        r = (u8)(cur | (r & bg));
        if (r != cur) { reach[p] = r; any = true; }
    }
    r = 0;
    for (int i = len - 1; i >= 0; --i) {
        const size_t p = off + i * step;
        const u8 bg = mask[p] ? 0 : 1, cur = reach[p];
        r = (u8)(cur | (r & bg));
        if (r != cur) { reach[p] = r; any = true; }
    }
    if (any) *changed = 1;
}
// mask |= background that no corner reaches
__global__ void k_mask_fill_holes(u8 *mask, const u8 *reach, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !mask[i] && !reach[i]) mask[i] = 1;
}

// ---- img *= mask, on the padded image: every padded pixel that numpy.pad copies from (sy, sx) takes that pixel's mask ---
__device__ __forceinline__ int imod(int i, int p)
{
    const int t = i % p;
    return t < 0 ? t + p : t;
}
__device__ __forceinline__ int pad_src(int i, int n, int mode)
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case B2S_PAD_REFLECT: { if (n == 1) return 0; const int p = 2 * (n - 1), t = imod(i, p); return t < n ? t : p - t; }
    case B2S_PAD_SYMMETRIC: { const int p = 2 * n, t = imod(i, p); return t < n ? t : p - 1 - t; }
    case B2S_PAD_WRAP: return imod(i, n);
    case B2S_PAD_EDGE: return i < 0 ? 0 : n - 1;
    default: return -1;          // constant and the computed modes: the pad area does not copy pixels
    }
}
__global__ void k_mask_apply(B2sImg img, const u8 *mask, int base_pad, int rows, int cols, int mode)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= img.cols) return;
    const size_t plane = blockIdx.z;
    const int sy = pad_src(y - base_pad, rows, mode), sx = pad_src(x - base_pad, cols, mode);
    if (sy < 0 || sx < 0) return;
    float *p = img.ptr + plane * img.plane_stride + (size_t)y * img.pitch + x;
    *p = *p * (float)mask[plane * (size_t)rows * cols + (size_t)sy * cols + sx];
}

}  // namespace

// mask: n_planes x rows x cols bytes holding the thresholded image on entry and get_img_mask's result on return;
// tmp, reach: scratch of the same size; d_flag: one device int; h_flag: page-locked host int
int b2s_launch_img_mask(unsigned char *mask, unsigned char *tmp, unsigned char *reach, int rows, int cols, int close_k, int open_k,
                        int *d_flag, int *h_flag, int n_planes, cudaStream_t s)
{
    const size_t smem = sizeof(int) * ((size_t)cols + 1);
    if (smem > 200 * 1024) return B2S_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_morph_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 gr(rows, n_planes), gc((cols + 127) / 128, n_planes);
    auto morph = [&](int k, int erode) {   // rows then columns; mask -> tmp -> mask
        const int lo = -(k / 2), hi = k - 1 - k / 2;
        k_morph_rows<<<gr, 256, smem, s>>>(mask, tmp, rows, cols, lo, hi, erode);
        k_morph_cols<<<gc, 128, 0, s>>>(tmp, mask, rows, cols, lo, hi, erode);
    };
    morph(close_k, 0); morph(close_k, 1);     // MORPH_CLOSE
    morph(open_k, 1); morph(open_k, 0);       // MORPH_OPEN
    k_flood_init<<<n_planes, 256, 0, s>>>(mask, reach, rows, cols);
    const dim3 fr((rows + 63) / 64, n_planes), fc((cols + 63) / 64, n_planes);
    for (int round = 0; round < 100000; ++round) {
        cudaMemsetAsync(d_flag, 0, sizeof(int), s);
        for (int k = 0; k < 2; ++k) {
            k_flood_sweep<<<fr, 64, 0, s>>>(mask, reach, rows, cols, 0, d_flag);
            k_flood_sweep<<<fc, 64, 0, s>>>(mask, reach, rows, cols, 1, d_flag);
        }
        cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return B2S_ERR_CUDA;
        if (!*h_flag) break;
    }
    const size_t n = (size_t)n_planes * rows * cols;
    k_mask_fill_holes<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mask, reach, n);
    return cudaGetLastError() == cudaSuccess ? B2S_OK : B2S_ERR_CUDA;
}

void b2s_launch_mask_threshold_f32(const B2sImg &padded, int base_pad, int rows, int cols, double thr, const double *thr_pp,
                                   unsigned char *mask, int n_planes, cudaStream_t s)
{
    k_mask_threshold_f32<<<dim3((cols + 255) / 256, rows, n_planes), 256, 0, s>>>(padded, base_pad, rows, cols, thr, thr_pp, mask);
}
void b2s_launch_mask_threshold_int(const void *in, int dtype, size_t n_per_plane, double thr, const double *thr_pp,
                                   unsigned char *mask, int n_planes, cudaStream_t s)
{
    const dim3 g((unsigned)((n_per_plane + 255) / 256), n_planes);
    if (dtype == B2S_U16) k_mask_threshold_int<<<g, 256, 0, s>>>((const unsigned short *)in, n_per_plane, thr, thr_pp, mask);
    else k_mask_threshold_int<<<g, 256, 0, s>>>((const unsigned char *)in, n_per_plane, thr, thr_pp, mask);
}
void b2s_launch_mask_apply(const B2sImg &padded, const unsigned char *mask, int base_pad, int rows, int cols, int pad_mode,
                           int n_planes, cudaStream_t s)
{
    k_mask_apply<<<dim3((padded.cols + 255) / 256, padded.rows, n_planes), 256, 0, s>>>(padded, mask, base_pad, rows, cols, pad_mode);
}
