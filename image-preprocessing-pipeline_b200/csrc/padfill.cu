// numpy.pad modes whose pad area is COMPUTED rather than copied: 'maximum', 'minimum', 'mean', 'median', 'linear_ramp'
// (and 'empty', filled with zeros).  filter_streaks accepts all eleven numpy modes (pystripe/core.py:1088-1110); the
// copying modes (reflect, wrap, symmetric, edge, constant) are index tables inside k_prologue, these five run after it
// on the padded log-domain image, in numpy's order: axis 0 first, over the original columns only, then axis 1 over EVERY
// row of the padded array (numpy/lib/_arraypad_impl.py: _view_roi, _get_stats, _get_linear_ramps).
//
// Rounding contract (float32 image, as filter_streaks pads log1p(img)):
//   mean, axis 0    sum of the column accumulated row by row (numpy reduces a non-contiguous axis with the reduced axis
//                   outermost), then sum / float32(n)
//   mean, axis 1    numpy's pairwise summation (blocks of <= 128 elements, 8 interleaved accumulators, halves split at a
//                   multiple of 8), then sum / float32(n)
//   median          the middle element, or (a + b) / 2 of the two middle elements, in float32
//   linear_ramp     numpy.linspace(0, edge, num=width, endpoint=False) evaluated in float64 (end_values is a numpy int64
//                   scalar): float32(i * (edge / width)); when ANY step of that side of the whole plane is zero numpy
//                   switches to float32((i / width) * edge) for the whole side — a per-plane, per-side flag here
// Not performance-critical: these modes exist for parity with the reference's argument surface.
#include <cstdio>

#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

struct Geo {
    float *p;       // plane 0 of the padded image
    size_t plane_stride;
    int pitch, PH, PW;
    int base, rows, cols;   // original area: rows [base, base + rows), columns [base, base + cols)
};

__device__ __forceinline__ float ramp_value(int i, int width, float edge, bool any_zero)
{
    const double d = (double)edge, n = (double)width, x = (double)i;
    const double y = any_zero ? __dmul_rn(__ddiv_rn(x, n), d) : __dmul_rn(x, __ddiv_rn(d, n));
    return (float)__dadd_rn(y, 0.0);
}

// ---- flags: any edge == 0 on a side of a plane (linear_ramp's `any_step_zero`) -------------------------------------
__global__ void k_ramp_flags(Geo g, unsigned *flags, int axis)
{
    float *pl = g.p + blockIdx.y * g.plane_stride;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (axis == 0) {
        if (idx >= g.cols) return;
        if (pl[(size_t)g.base * g.pitch + g.base + idx] == 0.f) atomicOr(flags + 4 * blockIdx.y + 0, 1u);
        if (pl[(size_t)(g.base + g.rows - 1) * g.pitch + g.base + idx] == 0.f) atomicOr(flags + 4 * blockIdx.y + 1, 1u);
    } else {
        if (idx >= g.PH) return;
        if (pl[(size_t)idx * g.pitch + g.base] == 0.f) atomicOr(flags + 4 * blockIdx.y + 2, 1u);
        if (pl[(size_t)idx * g.pitch + g.base + g.cols - 1] == 0.f) atomicOr(flags + 4 * blockIdx.y + 3, 1u);
    }
}

// ---- axis 0: one thread per original column -------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pad_axis0(Geo g, int mode, const unsigned *flags)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= g.cols) return;
    float *col = g.p + blockIdx.y * g.plane_stride + g.base + c;
    const int top = g.base, bottom = g.PH - g.base - g.rows;
    if (mode == B2S_PAD_LINEAR_RAMP) {
        const float e0 = col[(size_t)g.base * g.pitch], e1 = col[(size_t)(g.base + g.rows - 1) * g.pitch];
        const bool z0 = flags[4 * blockIdx.y + 0] != 0, z1 = flags[4 * blockIdx.y + 1] != 0;
        for (int i = 0; i < top; ++i) col[(size_t)i * g.pitch] = ramp_value(i, top, e0, z0);
        for (int i = 0; i < bottom; ++i) col[(size_t)(g.PH - 1 - i) * g.pitch] = ramp_value(i, bottom, e1, z1);
        return;
    }
    float s = col[(size_t)g.base * g.pitch];
    for (int r = 1; r < g.rows; ++r) {
        const float v = col[(size_t)(g.base + r) * g.pitch];
        if (mode == B2S_PAD_MAXIMUM) s = fmaxf(s, v);
        else if (mode == B2S_PAD_MINIMUM) s = fminf(s, v);
        else s = __fadd_rn(s, v);
    }
    if (mode == B2S_PAD_MEAN) s = __fdiv_rn(s, (float)g.rows);
    for (int i = 0; i < top; ++i) col[(size_t)i * g.pitch] = s;
    for (int i = 0; i < bottom; ++i) col[(size_t)(g.base + g.rows + i) * g.pitch] = s;
}

// numpy's pairwise float32 sum of a[0..n) (umath loops: FLOAT_pairwise_sum), recursion unrolled with an explicit stack
__device__ float pairwise_sum(const float *a, int n)
{
    int st_a[32], st_n[32], st_state[32];
    float st_acc[32];
    int sp = 0;
    st_a[0] = 0; st_n[0] = n; st_state[0] = 0;
    float ret = 0.f;
    while (sp >= 0) {
        const int o = st_a[sp], m = st_n[sp];
        if (st_state[sp] == 0) {
            if (m < 8) {
                float r = -0.0f;
                for (int i = 0; i < m; ++i) r = __fadd_rn(r, a[o + i]);
                ret = r; --sp;
            } else if (m <= 128) {
                float r[8];
                for (int j = 0; j < 8; ++j) r[j] = a[o + j];
                int i = 8;
                for (; i < m - (m % 8); i += 8)
                    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[o + i + j]);
                float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
                for (; i < m; ++i) res = __fadd_rn(res, a[o + i]);
                ret = res; --sp;
            } else {
                int n2 = m / 2;
                n2 -= n2 % 8;
                st_state[sp] = 1;
                ++sp;
                st_a[sp] = o; st_n[sp] = n2; st_state[sp] = 0;
            }
        } else if (st_state[sp] == 1) {      // left half done
            int n2 = m / 2;
            n2 -= n2 % 8;
            st_acc[sp] = ret;
            st_state[sp] = 2;
            ++sp;
            st_a[sp] = o + n2; st_n[sp] = m - n2; st_state[sp] = 0;
        } else {
            ret = __fadd_rn(st_acc[sp], ret);
            --sp;
        }
    }
    return ret;
}

// ---- axis 1: one CTA per row of the padded array ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pad_axis1(Geo g, int mode, const unsigned *flags)
{
    extern __shared__ float s_row[];
    __shared__ float s_red[4];
    __shared__ float s_stat;
    float *row = g.p + blockIdx.y * g.plane_stride + (size_t)blockIdx.x * g.pitch;
    const int left = g.base, right = g.PW - g.base - g.cols;
    const int tid = threadIdx.x;
    if (mode == B2S_PAD_LINEAR_RAMP) {
        const float e0 = row[g.base], e1 = row[g.base + g.cols - 1];
        const bool z0 = flags[4 * blockIdx.y + 2] != 0, z1 = flags[4 * blockIdx.y + 3] != 0;
        __syncthreads();
        for (int i = tid; i < left; i += 128) row[i] = ramp_value(i, left, e0, z0);
        for (int i = tid; i < right; i += 128) row[g.PW - 1 - i] = ramp_value(i, right, e1, z1);
        return;
    }
    if (mode == B2S_PAD_MEAN) {
        for (int x = tid; x < g.cols; x += 128) s_row[x] = row[g.base + x];
        __syncthreads();
        if (tid == 0) s_stat = __fdiv_rn(pairwise_sum(s_row, g.cols), (float)g.cols);
    } else {
        const bool mx = mode == B2S_PAD_MAXIMUM;
        float s = row[g.base];
        for (int x = tid; x < g.cols; x += 128) { const float v = row[g.base + x]; s = mx ? fmaxf(s, v) : fminf(s, v); }
        for (int o = 16; o; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, s, o); s = mx ? fmaxf(s, v) : fminf(s, v); }
        if ((tid & 31) == 0) s_red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float r = s_red[0];
            for (int w = 1; w < 4; ++w) r = mx ? fmaxf(r, s_red[w]) : fminf(r, s_red[w]);
            s_stat = r;
        }
    }
    __syncthreads();
    const float s = s_stat;
    for (int i = tid; i < left; i += 128) row[i] = s;
    for (int i = tid; i < right; i += 128) row[g.base + g.cols + i] = s;
}

// ---- median: one CTA per column (axis 0) or per row (axis 1); bitonic sort of the sequence in shared memory ---------
__global__ void __launch_bounds__(256) k_pad_median(Geo g, int axis, int npow2)
{
    extern __shared__ float s_v[];
    float *pl = g.p + blockIdx.y * g.plane_stride;
    const int n = axis == 0 ? g.rows : g.cols;
    const int tid = threadIdx.x;
    for (int i = tid; i < npow2; i += 256) {
        float v = __int_as_float(0x7f800000);     // +inf pads the sort
        if (i < n) v = axis == 0 ? pl[(size_t)(g.base + i) * g.pitch + g.base + blockIdx.x] : pl[(size_t)blockIdx.x * g.pitch + g.base + i];
        s_v[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const float a = s_v[i], b = s_v[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_v[i] = b; s_v[l] = a; }
                }
            }
            __syncthreads();
        }
    // numpy: mean of part[n/2 - 1 : n/2 + 1] (even) or of the single middle element (odd), float32
    const float med = (n & 1) ? s_v[n / 2] : __fdiv_rn(__fadd_rn(s_v[n / 2 - 1], s_v[n / 2]), 2.0f);
    if (axis == 0) {
        float *col = pl + g.base + blockIdx.x;
        const int top = g.base, bottom = g.PH - g.base - g.rows;
        for (int i = tid; i < top; i += 256) col[(size_t)i * g.pitch] = med;
        for (int i = tid; i < bottom; i += 256) col[(size_t)(g.base + g.rows + i) * g.pitch] = med;
    } else {
        float *row = pl + (size_t)blockIdx.x * g.pitch;
        const int left = g.base, right = g.PW - g.base - g.cols;
        for (int i = tid; i < left; i += 256) row[i] = med;
        for (int i = tid; i < right; i += 256) row[g.base + g.cols + i] = med;
    }
}

}  // namespace

// largest sequence the median sort holds in shared memory (floats, power of two)
int b2s_pad_fill_supported(int mode, int rows, int cols)
{
    if (mode != B2S_PAD_MEDIAN) return 1;
    return rows <= 32768 && cols <= 32768;
}

void b2s_launch_pad_fill(int mode, const B2sImg &img, int base_pad, int rows, int cols, unsigned *flags, int n_planes, cudaStream_t s)
{
    if (mode == B2S_PAD_EMPTY || n_planes <= 0) return;     // the prologue wrote zeros into the pad area
    Geo g{img.ptr, img.plane_stride, img.pitch, img.rows, img.cols, base_pad, rows, cols};
    if (mode == B2S_PAD_MEDIAN) {
        for (int axis = 0; axis < 2; ++axis) {
            const int n = axis == 0 ? rows : cols;
            int np2 = 1;
            while (np2 < n) np2 <<= 1;
            const size_t bytes = sizeof(float) * np2;
            cudaFuncSetAttribute(k_pad_median, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            k_pad_median<<<dim3(axis == 0 ? cols : g.PH, n_planes), 256, bytes, s>>>(g, axis, np2);
        }
        return;
    }
    if (mode == B2S_PAD_LINEAR_RAMP) {
        cudaMemsetAsync(flags, 0, sizeof(unsigned) * 4 * n_planes, s);
        k_ramp_flags<<<dim3((cols + 127) / 128, n_planes), 128, 0, s>>>(g, flags, 0);
    }
    k_pad_axis0<<<dim3((cols + 127) / 128, n_planes), 128, 0, s>>>(g, mode, flags);
    if (mode == B2S_PAD_LINEAR_RAMP) k_ramp_flags<<<dim3((g.PH + 127) / 128, n_planes), 128, 0, s>>>(g, flags, 1);
    const size_t bytes = mode == B2S_PAD_MEAN ? sizeof(float) * (size_t)cols : 0;
    if (bytes > 48 * 1024) cudaFuncSetAttribute(k_pad_axis1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_pad_axis1<<<dim3(g.PH, n_planes), 128, bytes, s>>>(g, mode, flags);
}
