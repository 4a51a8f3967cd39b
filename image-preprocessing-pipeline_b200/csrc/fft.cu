// Batched arbitrary-length real FFT -> Gaussian notch on the packed spectrum -> inverse FFT, entirely in shared memory.
//
// Replaces np_filter_coefficient (pystripe/core.py:749-754): scipy.fftpack.rfft along one axis of a detail sub-band,
// multiplication by np_notch (core.py:637-667) INDEXED BY PACKED ARRAY POSITION (r0, re1, im1, re2, im2, ...), irfft.
// Two real sequences ride one complex transform (z = a + i b); the spectra are separated, multiplied by the notch
// and recombined in a single pass between the forward and the inverse Stockham passes.  Lengths are arbitrary
// (sub-band sides such as 1333 = 31*43 or the prime 347): the length is factorised into radices 4,2,3,5,7 and
// whatever primes remain; every pass is a generic radix-R Stockham step.
//
// This stage is NOT bit-identical to pocketfft's float32 rounding (it is at least as accurate); see DESIGN.md.
#include "b2s_internal.h"

namespace {

constexpr int kNT = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

struct NotchArgs {
    B2sImg img;
    const float2 *tw;   // exp(-2 pi i k / n)
    const float *g;     // notch over packed positions
    int n, nseq, along_cols;
    int n_factors;
    int factors[32];
    int pairs_per_plane;
};

// one generic radix-R Stockham pass: thread per output element
template <bool INV>
__device__ __forceinline__ void stockham_pass(const float2 *__restrict__ in, float2 *__restrict__ out,
                                              const float2 *__restrict__ tw, int n, int R, int Ns)
{
    const int m = n / R;
    const int stride_tw = n / (Ns * R);
    for (int idx = threadIdx.x; idx < n; idx += kNT) {
        const int q = idx / m;
        const int j = idx - q * m;
        const int k = j % Ns;
        const int step = (k + q * Ns) * stride_tw;
        float2 acc = in[j];
        int e = step;
        for (int r = 1; r < R; ++r) {
            float2 w = tw[e];
            if (INV) w.y = -w.y;
            const float2 v = in[j + r * m];
            acc.x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc.x));
            acc.y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc.y));
            e += step;
            if (e >= n) e -= n;
        }
        out[(j - k) * R + k + q * Ns] = acc;
    }
}

__global__ void __launch_bounds__(kNT) k_notch(NotchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n;
    float2 *bufA = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufB = bufA + n;
    float2 *tw = bufB + n;
    float *g = reinterpret_cast<float *>(tw + n);

    for (int i = threadIdx.x; i < n; i += kNT) { tw[i] = a.tw[i]; g[i] = a.g[i]; }

    const float inv_n = 1.0f / (float)n;
    float *plane = a.img.ptr + (size_t)blockIdx.y * a.img.plane_stride;
    for (int pair = blockIdx.x; pair < a.pairs_per_plane; pair += gridDim.x) {
        const int s0 = 2 * pair, s1 = 2 * pair + 1;
        const bool has1 = s1 < a.nseq;
        __syncthreads();
        if (!a.along_cols) {
            const float *r0 = plane + (size_t)s0 * a.img.pitch;
            const float *r1 = plane + (size_t)s1 * a.img.pitch;
            for (int t = threadIdx.x; t < n; t += kNT) bufA[t] = make_float2(r0[t], has1 ? r1[t] : 0.f);
        } else {
            for (int t = threadIdx.x; t < n; t += kNT) {
                const float *p = plane + (size_t)t * a.img.pitch + s0;
                bufA[t] = make_float2(p[0], has1 ? p[1] : 0.f);
            }
        }
        __syncthreads();
        float2 *cur = bufA, *nxt = bufB;
        int Ns = 1;
        for (int f = 0; f < a.n_factors; ++f) {
            stockham_pass<false>(cur, nxt, tw, n, a.factors[f], Ns);
            Ns *= a.factors[f];
            __syncthreads();
            float2 *t = cur; cur = nxt; nxt = t;
        }
        // separate the two spectra, apply the notch on packed positions, recombine
        for (int k = threadIdx.x; k <= n / 2; k += kNT) {
            const int kp = k == 0 ? 0 : n - k;
            const float2 zk = cur[k], zp = cur[kp];
            float ar = 0.5f * (zk.x + zp.x), ai = 0.5f * (zk.y - zp.y);
            float br = 0.5f * (zk.y + zp.y), bi = -0.5f * (zk.x - zp.x);
            const float gr = k == 0 ? g[0] : g[2 * k - 1];
            const float gi = (k == 0 || 2 * k == n) ? 0.f : g[2 * k];
            ar *= gr; br *= gr; ai *= gi; bi *= gi;
            cur[k] = make_float2(ar - bi, ai + br);
            if (kp != k) cur[kp] = make_float2(ar + bi, br - ai);
        }
        __syncthreads();
        Ns = 1;
        for (int f = 0; f < a.n_factors; ++f) {
            stockham_pass<true>(cur, nxt, tw, n, a.factors[f], Ns);
            Ns *= a.factors[f];
            __syncthreads();
            float2 *t = cur; cur = nxt; nxt = t;
        }
        if (!a.along_cols) {
            float *r0 = plane + (size_t)s0 * a.img.pitch;
            float *r1 = plane + (size_t)s1 * a.img.pitch;
            for (int t = threadIdx.x; t < n; t += kNT) {
                const float2 z = cur[t];
                r0[t] = z.x * inv_n;
                if (has1) r1[t] = z.y * inv_n;
            }
        } else {
            for (int t = threadIdx.x; t < n; t += kNT) {
                const float2 z = cur[t];
                float *p = plane + (size_t)t * a.img.pitch + s0;
                p[0] = z.x * inv_n;
                if (has1) p[1] = z.y * inv_n;
            }
        }
    }
}

}  // namespace

size_t b2s_notch_smem(int n) { return (size_t)n * (3 * sizeof(float2) + sizeof(float)); }

void b2s_launch_notch(const B2sFftPlan &fp, const float *d_notch, const B2sImg &img, int along_cols, int n_planes,
                      int sm_count, cudaStream_t s)
{
    NotchArgs a;
    a.img = img;
    a.tw = fp.d_twiddle;
    a.g = d_notch;
    a.n = fp.n;
    a.along_cols = along_cols;
    a.nseq = along_cols ? img.cols : img.rows;
    a.n_factors = fp.n_factors;
    for (int i = 0; i < fp.n_factors; ++i) a.factors[i] = fp.factors[i];
    a.pairs_per_plane = (a.nseq + 1) / 2;
    const size_t bytes = b2s_notch_smem(fp.n);
    cudaFuncSetAttribute(k_notch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    int bx = a.pairs_per_plane;
    const int cap = (sm_count * 8 + n_planes - 1) / n_planes;
    if (bx > cap) bx = cap > 0 ? cap : 1;
    k_notch<<<dim3(bx, n_planes), kNT, bytes, s>>>(a);
}
