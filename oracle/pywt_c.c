/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under image-preprocessing-pipeline_b200/ may link or call this.
 *
 * CPU restatement of the third-party arithmetic the reference's hot path executes inside PyWavelets
 * (absent from /root/reference and from this machine, version unpinned by the reference — SURVEY.md §8c):
 *   pywt/_extensions/c/convolution.template.c : <T>_downsampling_convolution (MODE_SYMMETRIC),
 *                                               <T>_upsampling_convolution_valid_sf
 *   pywt/_extensions/c/wt.template.c          : <T>_dec_a/_dec_d/_idwt, dwt_max_level, dwt_buffer_length
 * called from /root/reference/pystripe/core.py:927 (wavedec2) and :939 (waverec2).
 * PARITY UNPINNED for this file: there is no PyWavelets here to compare with; what is pinned is the
 * perfect-reconstruction property, the published filter tables and the summation order documented below.
 *
 * Also: thin array wrappers over glibc's log1pf / expm1f.  The reference evaluates log1p/expm1 through
 * numexpr (core.py:184,194) which calls libm; numpy does the same unless it dispatches to AVX512 SVML
 * (checked: identical to libm with NPY_DISABLE_CPU_FEATURES=AVX512F...).  The oracle pins libm.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, like x86-64 baseline wheels).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* ---- pywt: size helpers (wt.template.c / common.c) ------------------------------------------- */
size_t orc_dwt_buffer_length(size_t input_len, size_t filter_len) { return (input_len + filter_len - 1) / 2; }
size_t orc_idwt_buffer_length(size_t coeffs_len, size_t filter_len) { return 2 * coeffs_len - filter_len + 2; }
static int size_log2(size_t x) { int r = -1; while (x) { ++r; x >>= 1; } return r; }
int orc_dwt_max_level(size_t input_len, size_t filter_len)
{
    if (filter_len <= 1 || input_len < (filter_len - 1)) return 0;
    return size_log2(input_len / (filter_len - 1));
}

/* ---- pywt: downsampling convolution, symmetric mode, step 2 ------------------------------------
 * out[o] = sum_j filter[j] * xext[i - j], i = 2*o + 1, xext = half-sample symmetric extension.
 * Summation order (this is what makes float32 results reproducible):
 *   left overhang and centre : j ascending 0..F-1
 *   right overhang (i >= N)  : first the reflected part with filter index i-N, i-N-1, ..., 0,
 *                              then the in-range part j = i-N+1 .. F-1
 *   F > N                    : left-reflected part, then in-range, then right-reflected (see below)
 * every term is a separate multiply followed by a separate add into the running sum (sum starts at 0).
 */
#define DEFINE_DOWNCONV(T, NAME)                                                                       \
void NAME(const T *input, size_t N, const T *filter, size_t F, T *output)                              \
{                                                                                                      \
    const size_t step = 2;                                                                             \
    size_t i = step - 1, o = 0;                                                                        \
    /* left boundary overhang */                                                                       \
    for (; i < F && i < N; i += step, ++o) {                                                           \
        T sum = 0;                                                                                     \
        size_t j;                                                                                      \
        for (j = 0; j <= i; ++j) sum += filter[j] * input[i - j];                                      \
        while (j < F) {                                                                                \
            size_t k;                                                                                  \
            for (k = 0; k < N && j < F; ++j, ++k) sum += filter[j] * input[k];                         \
            for (k = 0; k < N && j < F; ++k, ++j) sum += filter[j] * input[N - 1 - k];                 \
        }                                                                                              \
        output[o] = sum;                                                                               \
    }                                                                                                  \
    /* centre, N >= F */                                                                               \
    for (; i < N; i += step, ++o) {                                                                    \
        T sum = 0;                                                                                     \
        size_t j;                                                                                      \
        for (j = 0; j < F; ++j) sum += input[i - j] * filter[j];                                       \
        output[o] = sum;                                                                               \
    }                                                                                                  \
    /* centre, F > N : both ends overhang */                                                           \
    for (; i < F; i += step, ++o) {                                                                    \
        T sum = 0;                                                                                     \
        size_t j = 0;                                                                                  \
        while (i - j >= N) {                                                                           \
            size_t k;                                                                                  \
            for (k = 0; k < N && i - j >= N; ++j, ++k) sum += filter[i - N - j] * input[N - 1 - k];    \
            for (k = 0; k < N && i - j >= N; ++j, ++k) sum += filter[i - N - j] * input[k];            \
        }                                                                                              \
        for (; j <= i; ++j) sum += filter[j] * input[i - j];                                           \
        while (j < F) {                                                                                \
            size_t k;                                                                                  \
            for (k = 0; k < N && j < F; ++j, ++k) sum += filter[j] * input[k];                         \
            for (k = 0; k < N && j < F; ++k, ++j) sum += filter[j] * input[N - 1 - k];                 \
        }                                                                                              \
        output[o] = sum;                                                                               \
    }                                                                                                  \
    /* right boundary overhang */                                                                      \
    for (; i < N + F - 1; i += step, ++o) {                                                            \
        T sum = 0;                                                                                     \
        size_t j = 0;                                                                                  \
        while (i - j >= N) {                                                                           \
            size_t k;                                                                                  \
            for (k = 0; k < N && i - j >= N; ++j, ++k) sum += filter[i - N - j] * input[N - 1 - k];    \
            for (k = 0; k < N && i - j >= N; ++j, ++k) sum += filter[i - N - j] * input[k];            \
        }                                                                                              \
        for (; j < F; ++j) sum += filter[j] * input[i - j];                                            \
        output[o] = sum;                                                                               \
    }                                                                                                  \
}
DEFINE_DOWNCONV(float, orc_downconv_sym_f32)
DEFINE_DOWNCONV(double, orc_downconv_sym_f64)

/* ---- pywt: upsampling convolution, "valid" part, accumulating into output ---------------------- */
#define DEFINE_UPCONV(T, NAME)                                                                         \
int NAME(const T *input, size_t N, const T *filter, size_t F, T *output, size_t O)                     \
{                                                                                                      \
    if ((F % 2) || (N < F / 2)) return -1;                                                             \
    if (O != 2 * N - F + 2) return -2;                                                                 \
    size_t o, i;                                                                                       \
    for (o = 0, i = F / 2 - 1; i < N; ++i, o += 2) {                                                   \
        T sum_even = 0, sum_odd = 0;                                                                   \
        size_t j;                                                                                      \
        for (j = 0; j < F / 2; ++j) {                                                                  \
            sum_even += filter[j * 2] * input[i - j];                                                  \
            sum_odd += filter[j * 2 + 1] * input[i - j];                                               \
        }                                                                                              \
        output[o] += sum_even;                                                                         \
        output[o + 1] += sum_odd;                                                                      \
    }                                                                                                  \
    return 0;                                                                                          \
}
DEFINE_UPCONV(float, orc_upconv_valid_f32)
DEFINE_UPCONV(double, orc_upconv_valid_f64)

/* ---- axis drivers (pywt _dwt.pyx dwt_axis / idwt_axis copy strided lines into contiguous buffers) ---
 * 2-D only, row-major, in (ny, nx).  axis 0 = pywt axis -2 (rows index), axis 1 = pywt axis -1.
 */
#define DEFINE_DWT_AXIS(T, NAME, DOWN)                                                                 \
void NAME(const T *in, size_t ny, size_t nx, int axis, const T *dec_lo, const T *dec_hi, size_t F,     \
          T *out_a, T *out_d, T *scratch /* >= 3*max(ny,nx)+F */)                                      \
{                                                                                                      \
    if (axis == 1) {                                                                                   \
        size_t M = (nx + F - 1) / 2;                                                                   \
        for (size_t y = 0; y < ny; ++y) {                                                              \
            DOWN(in + y * nx, nx, dec_lo, F, out_a + y * M);                                           \
            DOWN(in + y * nx, nx, dec_hi, F, out_d + y * M);                                           \
        }                                                                                              \
    } else {                                                                                           \
        size_t M = (ny + F - 1) / 2;                                                                   \
        T *line = scratch, *oa = scratch + ny, *od = oa + M;                                           \
        for (size_t x = 0; x < nx; ++x) {                                                              \
            for (size_t y = 0; y < ny; ++y) line[y] = in[y * nx + x];                                  \
            DOWN(line, ny, dec_lo, F, oa);                                                             \
            DOWN(line, ny, dec_hi, F, od);                                                             \
            for (size_t y = 0; y < M; ++y) { out_a[y * nx + x] = oa[y]; out_d[y * nx + x] = od[y]; }   \
        }                                                                                              \
    }                                                                                                  \
}
DEFINE_DWT_AXIS(float, orc_dwt_axis_f32, orc_downconv_sym_f32)
DEFINE_DWT_AXIS(double, orc_dwt_axis_f64, orc_downconv_sym_f64)

/* idwt along one axis: out = 0; out += upconv(a, rec_lo); out += upconv(d, rec_hi)   (wt.template.c _idwt) */
#define DEFINE_IDWT_AXIS(T, NAME, UP)                                                                  \
int NAME(const T *ca, const T *cd, size_t ny, size_t nx, int axis, const T *rec_lo, const T *rec_hi,   \
         size_t F, T *out, T *scratch /* >= 4*max(ny,nx) */)                                           \
{                                                                                                      \
    int rc = 0;                                                                                        \
    if (axis == 1) {                                                                                   \
        size_t O = 2 * nx - F + 2;                                                                     \
        memset(out, 0, sizeof(T) * ny * O);                                                            \
        for (size_t y = 0; y < ny; ++y) {                                                              \
            rc |= UP(ca + y * nx, nx, rec_lo, F, out + y * O, O);                                      \
            rc |= UP(cd + y * nx, nx, rec_hi, F, out + y * O, O);                                      \
        }                                                                                              \
    } else {                                                                                           \
        size_t O = 2 * ny - F + 2;                                                                     \
        T *la = scratch, *ld = scratch + ny, *lo = ld + ny;                                            \
        for (size_t x = 0; x < nx; ++x) {                                                              \
            for (size_t y = 0; y < ny; ++y) { la[y] = ca[y * nx + x]; ld[y] = cd[y * nx + x]; }        \
            memset(lo, 0, sizeof(T) * O);                                                              \
            rc |= UP(la, ny, rec_lo, F, lo, O);                                                        \
            rc |= UP(ld, ny, rec_hi, F, lo, O);                                                        \
            for (size_t y = 0; y < O; ++y) out[y * nx + x] = lo[y];                                    \
        }                                                                                              \
    }                                                                                                  \
    return rc;                                                                                         \
}
DEFINE_IDWT_AXIS(float, orc_idwt_axis_f32, orc_upconv_valid_f32)
DEFINE_IDWT_AXIS(double, orc_idwt_axis_f64, orc_upconv_valid_f64)

/* ---- libm pins (reference: numexpr "log1p(img)" / "expm1(img)", core.py:184,194) ---------------- */
void orc_log1pf(const float *in, float *out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = log1pf(in[i]); }
void orc_expm1f(const float *in, float *out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = expm1f(in[i]); }
