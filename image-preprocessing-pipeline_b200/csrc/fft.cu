// Batched arbitrary-length real FFT -> Gaussian notch on the packed spectrum -> inverse FFT, entirely in shared memory.
//
// Replaces np_filter_coefficient (pystripe/core.py:749-754): scipy.fftpack.rfft along one axis of a detail sub-band,
// multiplication by np_notch (core.py:637-667) INDEXED BY PACKED ARRAY POSITION (r0, re1, im1, re2, im2, ...), irfft.
// Two real sequences ride one complex transform (z = a + i b); the spectra are separated, multiplied by the notch
// and recombined in a single pass between the forward and the inverse transform.
//
// Lengths are arbitrary (sub-band sides such as 1333 = 31*43, 676 = 4*13*13 or the prime 347), so the transform is a
// mixed-radix Stockham FFT whose passes follow pocketfft's repertoire:
//   * radix 4 and 2 butterflies;
//   * any odd radix R <= 43: one thread owns one butterfly; x_r and x_{R-r} are combined first (the cos part of the
//     DFT matrix only sees their sum, the sin part only their difference), which quarters the multiplications; the
//     (R-1)/2 x (R-1)/2 cos / sin tables live in constant memory and are read through the uniform data path, the
//     accumulators stay in registers, the arithmetic is packed FFMA2;
//   * larger prime factors (61, 101, 347, ...): the same sum/difference trick evaluated cooperatively, one thread per
//     output pair, operands staged in a third shared buffer.
// A CTA transforms G sequence pairs at once so that every pass has enough butterflies for all threads.
//
// This stage is NOT bit-identical to pocketfft's float32 rounding (it is at least as accurate); see DESIGN.md.
#include <cmath>
#include <vector>

#include "b2s_internal.h"

namespace {

constexpr int kFT = 128;            // threads per CTA
constexpr int kMaxRegH = 21;        // odd radices up to 2*21+1 = 43 keep their accumulators in registers

__host__ __device__ constexpr int pad4(int h) { return (h + 3) & ~3; }
__host__ __device__ constexpr int dft_table_offset(int h)   // floats before the table of radix 2h+1
{
    int o = 0;
    for (int k = 1; k < h; ++k) o += k * 2 * pad4(k);
    return o;
}
constexpr int kDftTableFloats = dft_table_offset(kMaxRegH + 1);

// radix 2h+1, row r-1 (r = 1..h): cos(2 pi q r / R) for q = 1..h (padded to a multiple of 4), then sin(2 pi q r / R)
__constant__ float c_dft[kDftTableFloats];

// asynchronous global -> shared copies: the whole gather of a group is in flight at once
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

struct NotchArgs {
    B2sImg img;
    const float2 *tw;     // exp(-2 pi i k / L), L = transform length (n, or n2 for Bluestein)
    const float *g;       // notch over packed positions
    const float2 *chirp;  // Bluestein: exp(-i pi k^2 / n), k < n
    const float2 *bf;     // Bluestein: FFT_L of the wrapped conjugate chirp, scaled by 1/L
    int n, nseq, along_cols;
    int L;                // transform length and per-sequence stride in shared memory
    int n_factors;        // factors of L
    int factors[32];
    int group;            // sequence pairs per CTA pass
    int has_large;        // a factor > 43 is present (third buffer allocated)
    int groups_per_plane;
};

// inter-stage twiddle exp(-+ 2 pi i e / n)
template <bool INV>
__device__ __forceinline__ float2 twiddle(const float2 *tw, int e)
{
    float2 w = tw[e];
    if (INV) w.y = -w.y;
    return w;
}

// ---- odd radix R = 2H+1, register butterflies -----------------------------------------------------------------
template <int H, bool INV>
__device__ __forceinline__ void pass_odd(const float2 *__restrict__ in, float2 *__restrict__ out,
                                         const float2 *__restrict__ tw, int n, int Ns, int G)
{
    constexpr int R = 2 * H + 1, HP = pad4(H), OFF = dft_table_offset(H);
    const int m = n / R;
    const int stride_tw = n / (Ns * R);
    for (int b = threadIdx.x; b < G * m; b += kFT) {
        const int g = b / m, j = b - g * m;
        const float2 *zin = in + g * n;
        float2 *zout = out + g * n;
        const int k = j % Ns;
        const int step = k * stride_tw;
        const float2 x0 = zin[j];
        float2 sum = x0;
        float2 A[H], B[H];
#pragma unroll
        for (int q = 0; q < H; ++q) { A[q] = make_float2(0.f, 0.f); B[q] = make_float2(0.f, 0.f); }
        int e1 = step, e2 = (int)(((long long)(R - 1) * step) % n);
#pragma unroll 1
        for (int r = 1; r <= H; ++r) {
            float2 a = zin[j + r * m], c = zin[j + (R - r) * m];
            if (Ns > 1) {
                a = cmul(a, twiddle<INV>(tw, e1));
                c = cmul(c, twiddle<INV>(tw, e2));
                e1 += step; if (e1 >= n) e1 -= n;
                e2 -= step; if (e2 < 0) e2 += n;
            }
            const float2 s = cadd(a, c), d = csub(a, c);
            sum = cadd(sum, s);
            const float *row = c_dft + OFF + (r - 1) * 2 * HP;
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float cq = row[q], sq = row[HP + q];
                A[q] = __ffma2_rn(s, make_float2(cq, cq), A[q]);
                B[q] = __ffma2_rn(d, make_float2(sq, sq), B[q]);
            }
        }
        const int obase = (j - k) * R + k;
        zout[obase] = sum;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            const float2 a = cadd(x0, A[q - 1]), bq = B[q - 1];
            // forward: X_q = x0 + A - iB, X_{R-q} = x0 + A + iB; inverse: the two swap
            const float2 lo = make_float2(a.x + bq.y, a.y - bq.x), hi = make_float2(a.x - bq.y, a.y + bq.x);
            zout[obase + q * Ns] = INV ? hi : lo;
            zout[obase + (R - q) * Ns] = INV ? lo : hi;
        }
    }
}

// ---- radix 4 / radix 2 ------------------------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ void pass_r4(const float2 *__restrict__ in, float2 *__restrict__ out,
                                        const float2 *__restrict__ tw, int n, int Ns, int G)
{
    const int m = n >> 2;
    const int stride_tw = n / (Ns * 4);
    for (int b = threadIdx.x; b < G * m; b += kFT) {
        const int g = b / m, j = b - g * m;
        const float2 *zin = in + g * n;
        float2 *zout = out + g * n;
        const int k = j % Ns;
        float2 a0 = zin[j], a1 = zin[j + m], a2 = zin[j + 2 * m], a3 = zin[j + 3 * m];
        if (Ns > 1) {
            const int step = k * stride_tw;
            a1 = cmul(a1, twiddle<INV>(tw, step));
            a2 = cmul(a2, twiddle<INV>(tw, 2 * step));
            a3 = cmul(a3, twiddle<INV>(tw, 3 * step));
        }
        const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
        // forward: multiply d13 by -i; inverse: by +i
        const float2 jd = INV ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        const int obase = (j - k) * 4 + k;
        zout[obase] = cadd(s02, s13);
        zout[obase + Ns] = cadd(d02, jd);
        zout[obase + 2 * Ns] = csub(s02, s13);
        zout[obase + 3 * Ns] = csub(d02, jd);
    }
}

template <bool INV>
__device__ __forceinline__ void pass_r2(const float2 *__restrict__ in, float2 *__restrict__ out,
                                        const float2 *__restrict__ tw, int n, int Ns, int G)
{
    const int m = n >> 1;
    const int stride_tw = n / (Ns * 2);
    for (int b = threadIdx.x; b < G * m; b += kFT) {
        const int g = b / m, j = b - g * m;
        const float2 *zin = in + g * n;
        float2 *zout = out + g * n;
        const int k = j % Ns;
        const float2 a0 = zin[j];
        float2 a1 = zin[j + m];
        if (Ns > 1) a1 = cmul(a1, twiddle<INV>(tw, k * stride_tw));
        const int obase = (j - k) * 2 + k;
        zout[obase] = cadd(a0, a1);
        zout[obase + Ns] = csub(a0, a1);
    }
}

// ---- large odd radix (prime factors > 43): cooperative evaluation --------------------------------------------
// phase 1: S[r][j] = x_r + x_{R-r}, D[r][j] = x_r - x_{R-r} (twiddled) into `tmp`; phase 2: one thread per (q, j) pair of
// outputs.  cos/sin(2 pi t / R) = tw[t * n / R].
template <bool INV>
__device__ __forceinline__ void pass_large(const float2 *__restrict__ in, float2 *__restrict__ out, float2 *__restrict__ tmp,
                                           const float2 *__restrict__ tw, int n, int R, int Ns, int G)
{
    const int H = (R - 1) >> 1;
    const int m = n / R;
    const int stride_tw = n / (Ns * R);
    const int per = H * m;   // (r, j) items per sequence
    for (int b = threadIdx.x; b < G * per; b += kFT) {
        const int g = b / per, rem = b - g * per;
        const int r = rem / m + 1, j = rem - (r - 1) * m;
        const float2 *zin = in + g * n;
        float2 a = zin[j + r * m], c = zin[j + (R - r) * m];
        if (Ns > 1) {
            const int k = j % Ns;
            const long long step = (long long)k * stride_tw;
            a = cmul(a, twiddle<INV>(tw, (int)((step * r) % n)));
            c = cmul(c, twiddle<INV>(tw, (int)((step * (R - r)) % n)));
        }
        float2 *t = tmp + g * n;
        t[(r - 1) * m + j] = cadd(a, c);
        t[(H + r - 1) * m + j] = csub(a, c);
    }
    __syncthreads();
    const int cstride = n / R;
    const int per2 = (H + 1) * m;   // q = 0..H
    for (int b = threadIdx.x; b < G * per2; b += kFT) {
        const int g = b / per2, rem = b - g * per2;
        const int q = rem / m, j = rem - q * m;
        const float2 *zin = in + g * n;
        const float2 *t = tmp + g * n;
        float2 *zout = out + g * n;
        const int k = j % Ns;
        const int obase = (j - k) * R + k;
        const float2 x0 = zin[j];
        float2 A = make_float2(0.f, 0.f), B = make_float2(0.f, 0.f);
        int idx = 0;
        for (int r = 1; r <= H; ++r) {
            idx += q; if (idx >= R) idx -= R;
            const float2 w = tw[idx * cstride];   // (cos, -sin)
            const float2 s = t[(r - 1) * m + j], d = t[(H + r - 1) * m + j];
            A = __ffma2_rn(s, make_float2(w.x, w.x), A);
            B = __ffma2_rn(d, make_float2(-w.y, -w.y), B);
        }
        const float2 a = cadd(x0, A);
        if (q == 0) { zout[obase] = a; continue; }
        const float2 lo = make_float2(a.x + B.y, a.y - B.x), hi = make_float2(a.x - B.y, a.y + B.x);
        zout[obase + q * Ns] = INV ? hi : lo;
        zout[obase + (R - q) * Ns] = INV ? lo : hi;
    }
}

template <bool INV>
__device__ __forceinline__ void fft_pass(int R, const float2 *in, float2 *out, float2 *tmp, const float2 *tw, int n, int Ns,
                                         int G)
{
    switch (R) {
    case 2: pass_r2<INV>(in, out, tw, n, Ns, G); break;
    case 4: pass_r4<INV>(in, out, tw, n, Ns, G); break;
#define X(HH) case 2 * HH + 1: pass_odd<HH, INV>(in, out, tw, n, Ns, G); break;
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21)
#undef X
    default: pass_large<INV>(in, out, tmp, tw, n, R, Ns, G); break;
    }
}

// all passes of one length-L transform over `ng` sequences; result ends in `cur`
template <bool INV>
__device__ __forceinline__ void run_passes(const NotchArgs &a, float2 *&cur, float2 *&nxt, float2 *tmp, const float2 *tw, int ng)
{
    int Ns = 1;
    for (int f = 0; f < a.n_factors; ++f) {
        const int R = a.factors[f];
        fft_pass<INV>(R, cur, nxt, tmp, tw, a.L, Ns, ng);
        Ns *= R;
        __syncthreads();
        float2 *t = cur; cur = nxt; nxt = t;
    }
}

// unnormalised forward DFT of length n of every sequence in `cur` (stride L).  Direct when L == n, else Bluestein:
// X_j = c_j * sum_k (z_k c_k) conj(c)_{j-k}, the convolution done cyclically at length L >= 2n-1.
__device__ __forceinline__ void dft_n(const NotchArgs &a, float2 *&cur, float2 *&nxt, float2 *tmp, const float2 *tw, int ng)
{
    const int n = a.n, L = a.L;
    if (L == n) { run_passes<false>(a, cur, nxt, tmp, tw, ng); return; }
    for (int i = threadIdx.x; i < ng * L; i += kFT) {
        const int gi = i / L, k = i - gi * L;
        float2 *z = cur + gi * L;
        z[k] = k < n ? cmul(z[k], __ldg(a.chirp + k)) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    run_passes<false>(a, cur, nxt, tmp, tw, ng);
    for (int i = threadIdx.x; i < ng * L; i += kFT) {
        const int gi = i / L, k = i - gi * L;
        float2 *z = cur + gi * L;
        z[k] = cmul(z[k], __ldg(a.bf + k));
    }
    __syncthreads();
    run_passes<true>(a, cur, nxt, tmp, tw, ng);
    for (int i = threadIdx.x; i < ng * n; i += kFT) {
        const int gi = i / n, k = i - gi * n;
        float2 *z = cur + gi * L;
        z[k] = cmul(z[k], __ldg(a.chirp + k));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kFT) k_notch(const NotchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, L = a.L, G = a.group;
    float2 *bufA = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufB = bufA + G * L;
    float2 *tmp = bufB + G * L;
    float2 *tw = tmp + (a.has_large ? G * L : 0);
    float *g = reinterpret_cast<float *>(tw + L);

    for (int i = threadIdx.x; i < L; i += kFT) tw[i] = a.tw[i];
    for (int i = threadIdx.x; i < n; i += kFT) g[i] = a.g[i];

    const float inv_n = 1.0f / (float)n;
    float *plane = a.img.ptr + (size_t)blockIdx.y * a.img.plane_stride;
    const int pairs = (a.nseq + 1) >> 1;
    for (int grp = blockIdx.x; grp < a.groups_per_plane; grp += gridDim.x) {
        const int pair0 = grp * G;
        const int ng = min(G, pairs - pair0);
        __syncthreads();
        // ---- gather: z[t] = seq(2p)[t] + i seq(2p+1)[t]
        if (!a.along_cols) {
            for (int i = threadIdx.x; i < ng * n; i += kFT) {
                const int gi = i / n, t = i - gi * n;
                const int s0 = 2 * (pair0 + gi);
                const float *r0 = plane + (size_t)s0 * a.img.pitch;
                float2 *z = bufA + gi * L + t;
                cp_async4(&z->x, r0 + t);
                if (s0 + 1 < a.nseq) cp_async4(&z->y, r0 + a.img.pitch + t);
                else z->y = 0.f;
            }
        } else {
            for (int i = threadIdx.x; i < ng * n; i += kFT) {
                const int t = i / ng, gi = i - t * ng;   // neighbouring threads read neighbouring columns
                const int s0 = 2 * (pair0 + gi);
                const float *p = plane + (size_t)t * a.img.pitch + s0;
                float2 *z = bufA + gi * L + t;
                if (s0 + 1 < a.nseq) cp_async8(z, p);
                else { cp_async4(&z->x, p); z->y = 0.f; }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        float2 *cur = bufA, *nxt = bufB;
        dft_n(a, cur, nxt, tmp, tw, ng);
        // ---- separate the two spectra, apply the notch on packed positions, recombine; the inverse transform is
        //      conj(DFT(conj(.))), so the recombined spectrum is stored conjugated
        const int half = n / 2 + 1;
        for (int i = threadIdx.x; i < ng * half; i += kFT) {
            const int gi = i / half, k = i - gi * half;
            float2 *z = cur + gi * L;
            const int kp = k == 0 ? 0 : n - k;
            const float2 zk = z[k], zp = z[kp];
            float ar = 0.5f * (zk.x + zp.x), ai = 0.5f * (zk.y - zp.y);
            float br = 0.5f * (zk.y + zp.y), bi = -0.5f * (zk.x - zp.x);
            const float gr = k == 0 ? g[0] : g[2 * k - 1];
            const float gim = (k == 0 || 2 * k == n) ? 0.f : g[2 * k];
            ar *= gr; br *= gr; ai *= gim; bi *= gim;
            z[k] = make_float2(ar - bi, -(ai + br));
            if (kp != k) z[kp] = make_float2(ar + bi, -(br - ai));
        }
        __syncthreads();
        dft_n(a, cur, nxt, tmp, tw, ng);
        // ---- scatter (undo the conjugation: only the sign of the imaginary part, i.e. of the second sequence)
        if (!a.along_cols) {
            for (int i = threadIdx.x; i < ng * n; i += kFT) {
                const int gi = i / n, t = i - gi * n;
                const int s0 = 2 * (pair0 + gi);
                float *r0 = plane + (size_t)s0 * a.img.pitch;
                const float2 z = cur[gi * L + t];
                r0[t] = z.x * inv_n;
                if (s0 + 1 < a.nseq) r0[a.img.pitch + t] = -z.y * inv_n;
            }
        } else {
            for (int i = threadIdx.x; i < ng * n; i += kFT) {
                const int t = i / ng, gi = i - t * ng;
                const int s0 = 2 * (pair0 + gi);
                float *p = plane + (size_t)t * a.img.pitch + s0;
                const float2 z = cur[gi * L + t];
                p[0] = z.x * inv_n;
                if (s0 + 1 < a.nseq) p[1] = -z.y * inv_n;
            }
        }
    }
}

void upload_dft_tables()
{
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    std::vector<float> t(kDftTableFloats, 0.f);
    for (int h = 1; h <= kMaxRegH; ++h) {
        const int R = 2 * h + 1, hp = pad4(h), off = dft_table_offset(h);
        for (int r = 1; r <= h; ++r)
            for (int q = 1; q <= h; ++q) {
                const double ang = 2.0 * M_PI * (double)((q * r) % R) / (double)R;
                t[off + (r - 1) * 2 * hp + (q - 1)] = (float)std::cos(ang);
                t[off + (r - 1) * 2 * hp + hp + (q - 1)] = (float)std::sin(ang);
            }
    }
    cudaMemcpyToSymbol(c_dft, t.data(), sizeof(float) * kDftTableFloats);
    done[dev] = true;
}

size_t smem_for(int n, int L, int group, int has_large)
{
    return (size_t)L * ((size_t)(has_large ? 3 : 2) * group * sizeof(float2) + sizeof(float2)) + (size_t)n * sizeof(float);
}

int largest_prime_factor(int n)
{
    int best = 1;
    for (int p = 2; (long long)p * p <= n; ++p)
        while (n % p == 0) { best = p; n /= p; }
    return n > 1 ? n : best;
}

// factors of L in pass order: 4s, 2s, then odd primes ascending
int factorize(int L, int *factors, int *has_large)
{
    int nf = 0, rem = L;
    *has_large = 0;
    while (rem % 4 == 0) { factors[nf++] = 4; rem /= 4; }
    while (rem % 2 == 0) { factors[nf++] = 2; rem /= 2; }
    for (int p = 3; rem > 1; p += 2) {
        if ((long long)p * p > rem) p = rem;
        while (rem % p == 0) {
            factors[nf++] = p;
            if (p > 2 * kMaxRegH + 1) *has_large = 1;
            rem /= p;
        }
    }
    return nf;
}

}  // namespace

// pocketfft switches to Bluestein when a length has a large prime factor; so do we (threshold: a prime factor > 127;
// primes 47..127 use the cooperative direct pass).  The convolution length is the smallest 13-smooth L >= 2n-1.
void b2s_fft_plan_init(B2sFftPlan *fp, int n, B2sFftHostTables *host)
{
    fp->n = n;
    fp->L = n;
    fp->bluestein = 0;
    if (largest_prime_factor(n) > 127) {
        int L = 2 * n - 1;
        while (largest_prime_factor(L) > 13) ++L;
        fp->L = L;
        fp->bluestein = 1;
    }
    const int L = fp->L;
    fp->n_factors = factorize(L, fp->factors, &fp->has_large);
    // sequence pairs per CTA: as many as fit in ~110 KB (two CTAs per SM), at most 8
    int g = 8;
    while (g > 1 && smem_for(n, L, g, fp->has_large) > 110 * 1024) --g;
    fp->group = g;
    if (!host) return;
    host->tw.resize(L);
    for (int k = 0; k < L; ++k) {
        const double ang = -2.0 * M_PI * (double)k / (double)L;
        host->tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    host->chirp.clear();
    host->bf.clear();
    if (fp->bluestein) {
        // c_k = exp(-i pi k^2 / n); k^2 is reduced modulo 2n so the angle keeps full precision
        std::vector<double> cr(n), ci(n);
        host->chirp.resize(n);
        for (int k = 0; k < n; ++k) {
            const long long k2 = ((long long)k * k) % (2LL * n);
            const double ang = -M_PI * (double)k2 / (double)n;
            cr[k] = std::cos(ang); ci[k] = std::sin(ang);
            host->chirp[k] = make_float2((float)cr[k], (float)ci[k]);
        }
        // B[t] = conj(c)_{|t|} wrapped to length L; BF = DFT_L(B) / L in double (O(L^2) once per plan is too slow for
        // L ~ 3000, so use the recurrence-free split: direct sum over the 2n-1 non-zero taps)
        std::vector<double> br(L, 0.0), bi(L, 0.0);
        for (int t = 0; t < n; ++t) {
            br[t] = cr[t]; bi[t] = -ci[t];
            if (t) { br[L - t] = cr[t]; bi[L - t] = -ci[t]; }
        }
        host->bf.resize(L);
        std::vector<double> wr(L), wi(L);
        for (int k = 0; k < L; ++k) { const double ang = -2.0 * M_PI * (double)k / (double)L; wr[k] = std::cos(ang); wi[k] = std::sin(ang); }
        for (int j = 0; j < L; ++j) {
            double sr = 0.0, si = 0.0;
            for (int t = 0; t < L; ++t) {
                if (br[t] == 0.0 && bi[t] == 0.0) continue;
                const int e = (int)(((long long)j * t) % L);
                sr += br[t] * wr[e] - bi[t] * wi[e];
                si += br[t] * wi[e] + bi[t] * wr[e];
            }
            host->bf[j] = make_float2((float)(sr / L), (float)(si / L));
        }
    }
}

size_t b2s_notch_smem(int n)
{
    B2sFftPlan fp;
    b2s_fft_plan_init(&fp, n, nullptr);
    return smem_for(n, fp.L, fp.group, fp.has_large);
}

void b2s_launch_notch(const B2sFftPlan &fp, const float *d_notch, const B2sImg &img, int along_cols, int n_planes,
                      int sm_count, cudaStream_t s)
{
    upload_dft_tables();
    NotchArgs a;
    a.img = img;
    a.tw = fp.d_twiddle;
    a.g = d_notch;
    a.chirp = fp.d_chirp;
    a.bf = fp.d_bf;
    a.n = fp.n;
    a.L = fp.L;
    a.along_cols = along_cols;
    a.nseq = along_cols ? img.cols : img.rows;
    a.n_factors = fp.n_factors;
    for (int i = 0; i < fp.n_factors; ++i) a.factors[i] = fp.factors[i];
    a.group = fp.group;
    a.has_large = fp.has_large;
    const int pairs = (a.nseq + 1) / 2;
    a.groups_per_plane = (pairs + fp.group - 1) / fp.group;
    const size_t bytes = smem_for(fp.n, fp.L, fp.group, fp.has_large);
    cudaFuncSetAttribute(k_notch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    int bx = a.groups_per_plane;
    const int cap = (sm_count * 4 + n_planes - 1) / n_planes;
    if (bx > cap) bx = cap > 0 ? cap : 1;
    k_notch<<<dim3(bx, n_planes), kFT, bytes, s>>>(a);
}
