// Internal declarations shared by the kernel translation units and api.cu.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#define B2S_MAX_TAPS 128
#define B2S_MAX_LEVELS 32

// Filter bank, float32 copies of the float64 tables (pywt casts its double tables to float for float32 data).
struct B2sTaps {
    int F;
    float dec_lo[B2S_MAX_TAPS], dec_hi[B2S_MAX_TAPS], rec_lo[B2S_MAX_TAPS], rec_hi[B2S_MAX_TAPS];
};

// One 2-D float32 image per plane inside a batched workspace buffer.
struct B2sImg {
    float *ptr;           // plane 0
    size_t plane_stride;  // floats between planes
    int pitch;            // floats between rows (multiple of 4)
    int rows, cols;       // logical size
};

// pointwise.cu ------------------------------------------------------------------------------------------------
struct B2sPrologueArgs {
    const void *in;        // u8/u16/f32 planes, contiguous (src_rows x src_cols)
    int in_dtype;
    int src_rows, src_cols;
    const float *flat;     // optional (src_rows x src_cols)
    int pad_mode, base_pad;
    int use_log1p;
    float pad_value;       // constant padding: the fill value (core.py:1101-1105), 0 otherwise
    const float *pad_value_pp; // optional: one fill value per plane (clip_min from multi-Otsu, core.py:1066-1077, 1101-1105)
    B2sImg out;            // padded image
    // numpy.pad as tables (built once per plan): padded rows grouped by the source row they replicate
    int n_groups;          // source rows + padded rows without a source (constant mode)
    const int *row_src;    // [n_groups] source row, -1 => zero rows
    const int *row_start;  // [n_groups + 1] CSR offsets into row_targets
    const int *row_targets;// padded row indices
    const int *colmap;     // [out.pitch] source column of each padded column, -1 => 0
    const float *lut;      // optional: log1p of every integer pixel value (no flat, integer input)
    unsigned *minmax;      // optional: per-plane {min key, ~max key} of the raw input, accumulated here (uniform check fused)
};
void b2s_launch_prologue(const B2sPrologueArgs &a, int n_planes, cudaStream_t s);

struct B2sEpilogueArgs {
    B2sImg in;             // padded reconstruction (log domain when use_log1p); or unused when !destripe
    const void *raw;       // when destripe is skipped: the (pre-processed) source image, raw_dtype
    int raw_dtype;
    int destripe;          // 1: take `in`, crop base_pad;   0: take `raw`
    int base_pad;
    int rows, cols;        // work image size (before rotation)
    int use_log1p;
    int int_path;          // 1: rint+clip to the integer work dtype right after expm1 (filter_streaks on integer input)
    int work_dtype;        // dtype the reference's array has at this point when int_path (U8/U16)
    double dark;
    const float *ls_sub;   // optional lightsheet subtrahend (rows x cols, float) — already min(ls, bg*w)
    int final_mode;        // 0: clip+trunc to out_dtype (integer d_type); 1: convert_to_16bit; 2: convert_to_8bit; 3: float32 out
    int shift;
    int out_dtype;
    int flip, rot;         // rot in {0,1,2,3} quarter turns (numpy.rot90 k)
    int f32_exact;         // every value of the final conversion is exact in float32 (no dark, or integral dark on integers)
    const unsigned *uniform_mm; // optional per-plane {min key, ~max key}: equal => all pixels equal => write zeros (core.py:1232)
    void *out;
    int out_rows, out_cols;
    const float *epi_thr;  // optional: thresholds of the plan's log value -> integer map (b2s_launch_epi_thresholds); fast epilogue only
    int epi_kmax;          // the largest integer that map produces
};
void b2s_launch_epilogue(const B2sEpilogueArgs &a, int n_planes, cudaStream_t s);
// the fast epilogue's map from a log value to an integer in [0, 65535]: expm1 -> [rint, clip to the work dtype] -> dark -> clip
struct B2sEpiFn { int int_path; float darkf; float hi_w; };
// thr[k] (65 536 floats) = the smallest float v that map sends to k or above (thr[0] = -inf, NaN above the largest value,
// which lands in *d_kmax): the epilogue finds the integer from an approximate exponential and two table entries
void b2s_launch_epi_thresholds(float *thr, const B2sEpiFn &f, int *d_kmax, cudaStream_t s);
// exhaustive check over float bit patterns [first, first + count): table lookup == direct evaluation; mismatches are counted
void b2s_launch_epi_table_check(const float *thr, const B2sEpiFn &f, int kmax, unsigned long long first, unsigned long long count,
                                unsigned long long *mismatches, cudaStream_t s);

// new_size (skimage.transform.resize, order 1, core.py:1356-1359) fused with the final conversion / orientation
struct B2sResizeArgs {
    const void *src;       // image after dark / lightsheet, `dtype`, rows x cols per plane, contiguous
    int dtype, rows, cols;
    int new_rows, new_cols;
    const int *iy0, *iy1, *ix0, *ix1;            // per output row / column: the two source indices
    const double *wy0, *wy1, *wx0, *wx1;         // and their float64 weights
    const unsigned *mm;    // per-plane {min key, ~max key} of the unfiltered image (the clip range)
    int mm_dtype;          // dtype the keys were taken in (src may be the anti-aliased float64 / float32 copy)
};
#define B2S_F64_INTERNAL 100   // resize source after the anti-aliasing Gaussian on an integer image (float64 in skimage)
// scipy.ndimage.gaussian_filter1d along one axis (mode='mirror', float64 accumulation in correlate1d's symmetric
// order): in (u8/u16/f32/f64) -> out (f64 when out_f64, else f32); w: 2r+1 weights on the device
void b2s_launch_gauss_aa(const void *in, int in_dtype, void *out, int out_f64, int rows, int cols, int axis,
                         const double *w, int radius, int n_planes, cudaStream_t s);
void b2s_resize_axis_table(int n_in, int n_out, int *idx0, int *idx1, double *w0, double *w1);   // host
void b2s_launch_resize_final(const B2sResizeArgs &r, const B2sEpilogueArgs &a, int n_planes, cudaStream_t s);

// per-plane "all pixels equal" flags (process_img core.py:1232)
// mm[2p] = min key, mm[2p+1] = ~max key (both initialised to 0xffffffff by a memset); standalone pass over the input
void b2s_launch_minmax(const void *in, int dtype, size_t plane_elems, int n_planes, unsigned *mm, cudaStream_t s);

// pre-processing ahead of the destripe: flat division, 5x5 Gaussian, block reduce.  in/out dtype per stage.
void b2s_launch_flat_divide(const void *in, int in_dtype, const float *flat, float *out, size_t plane_elems,
                            int n_planes, cudaStream_t s);
void b2s_launch_gauss5_u16(const uint16_t *in, uint16_t *out, int rows, int cols, int n_planes, cudaStream_t s);
void b2s_launch_gauss5_f32(const float *in, float *out, int rows, int cols, int n_planes, cudaStream_t s);
void b2s_launch_block_reduce(const void *in, int dtype, int rows, int cols, int by, int bx, int method, void *out,
                             int out_dtype, int out_rows, int out_cols, int n_planes, cudaStream_t s);
// isotropic down-sampling helpers (parallel_image_processor.py:417-433)
void b2s_launch_z_pair(const float *a, const float *b, int method, float *out, int64_t n, cudaStream_t s);
void b2s_launch_convert_f32(const float *in, int64_t n, int mode, int shift, void *out, cudaStream_t s);
void b2s_launch_math(int which, const float *in, float *out, int64_t n, cudaStream_t s);
void b2s_launch_log1p_lut(float *lut, int n, cudaStream_t s);

// padfill.cu -----------------------------------------------------------------------------------------------------
// numpy.pad modes that compute their pad area (linear_ramp, maximum, mean, median, minimum; empty = zeros): run on the
// padded image after the prologue wrote the interior and zeros around it.  flags: 4 unsigned per plane (scratch).
int b2s_pad_fill_supported(int mode, int rows, int cols);
void b2s_launch_pad_fill(int mode, const B2sImg &padded, int base_pad, int rows, int cols, unsigned *flags, int n_planes,
                         cudaStream_t s);

// f64path.cu ------------------------------------------------------------------------------------------------------
// integer pixels with log1p_normalization_needed=False: the reference runs pad -> wavedec2 -> notch -> waverec2 in float64
struct B2sF64Args {
    const void *in;                // work-size integer planes (u8 / u16), contiguous
    int in_dtype, rows, cols;
    int pad_mode, base_pad;
    double pad_value;
    int PH, PW, levels;
    int my[B2S_MAX_LEVELS + 1], mx[B2S_MAX_LEVELS + 1];
    int F;
    const double *dec_lo, *dec_hi, *rec_lo, *rec_hi;   // host pointers, F doubles each
    int n_passes, bidirectional;
    const double *notch[2][B2S_MAX_LEVELS + 1][2];      // device: response matrices of irfft(rfft(x) * g), n x n, [k][c]
    double *work;                  // b2s_f64_workspace_doubles() doubles per plane
    size_t plane_doubles;
    B2sImg out;                    // float32 padded image: rint / clip of the reconstruction (the common epilogue reads it)
    const unsigned char *mask;     // optional get_img_mask result (rows x cols bytes per plane): img *= mask ahead of the padding
};
size_t b2s_f64_workspace_doubles(int PH, int PW, int levels, const int *my, const int *mx);
void b2s_launch_f64_destripe(const B2sF64Args &a, int n_planes, cudaStream_t s);

// mask.cu ----------------------------------------------------------------------------------------------------------
// get_img_mask (core.py:475-489): `mask` holds img > threshold on entry, the closed / opened / hole-filled mask on return
int b2s_launch_img_mask(unsigned char *mask, unsigned char *tmp, unsigned char *reach, int rows, int cols, int close_k, int open_k,
                        int *d_flag, int *h_flag, int n_planes, cudaStream_t s);
void b2s_launch_mask_threshold_f32(const B2sImg &padded, int base_pad, int rows, int cols, double thr, const double *thr_pp,
                                   unsigned char *mask, int n_planes, cudaStream_t s);
void b2s_launch_mask_threshold_int(const void *in, int dtype, size_t n_per_plane, double thr, const double *thr_pp,
                                   unsigned char *mask, int n_planes, cudaStream_t s);
void b2s_launch_mask_apply(const B2sImg &padded, const unsigned char *mask, int base_pad, int rows, int cols, int pad_mode,
                           int n_planes, cudaStream_t s);

// deflate.cu -------------------------------------------------------------------------------------------------------
// one zlib stream (literal-only dynamic Huffman block) per strip of rows_per_strip rows, packed back to back into `out`
size_t b2s_deflate_bound_bytes(size_t plane_bytes, int strips_per_plane, int n_planes);
void b2s_launch_deflate(const void *in, size_t plane_bytes, size_t row_bytes, int rows, int rows_per_strip, int n_planes, unsigned *tmp,
                        unsigned *sizes, unsigned long long *offsets, void *out, size_t capacity, int *overflow, cudaStream_t s);

// stats.cu ---------------------------------------------------------------------------------------------------------
// exact intensity histogram of uint8 / uint16 planes, ADDED into `hist` (65 536 uint64 counters, per plane or one for all)
void b2s_launch_histogram(const void *in, int dtype, size_t plane_elems, int n_planes, unsigned long long *hist, int per_plane,
                          cudaStream_t s);

// dwt.cu --------------------------------------------------------------------------------------------------------
// forward level: in (ny x nx) -> cA,cH,cV,cD ((ny+F-1)/2 x (nx+F-1)/2); exact!=0 => reference summation order, no FMA
void b2s_launch_dwt_fwd(const B2sTaps &t, const B2sImg &in, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV,
                        const B2sImg &cD, int n_planes, int exact, int sm_count, cudaStream_t s, float *scratch = nullptr,
                        size_t scratch_plane_stride = 0);
// inverse level: sub-bands (my x mx; cA is read with that logical size) -> out (first out.rows x out.cols of the
// 2*my-F+2 x 2*mx-F+2 reconstruction)
void b2s_launch_dwt_inv(const B2sTaps &t, const B2sImg &cA, const B2sImg &cH, const B2sImg &cV, const B2sImg &cD,
                        const B2sImg &out, int n_planes, int exact, int sm_count, cudaStream_t s, float *scratch = nullptr,
                        size_t scratch_plane_stride = 0);
int b2s_dwt_max_smem(int F);
// long filters (F >= 42) run one kernel per axis through a scratch buffer: floats per plane for a level whose input is
// ny x nx (0: the fused kernels are used); the level-1 figure covers every level
size_t b2s_dwt_scratch_floats(int F, int ny, int nx);

// fft.cu --------------------------------------------------------------------------------------------------------
struct B2sFftPlan {      // per sequence length n
    int n;
    int L;               // transform length: n, or the Bluestein convolution length (>= 2n-1, 13-smooth)
    int bluestein;
    int n_factors;       // factors of L
    int factors[32];
    int group;           // sequence pairs transformed together by one CTA
    int has_large;       // a prime factor > 43 is present
    float2 *d_twiddle;   // L entries exp(-2 pi i k / L)
    float2 *d_chirp;     // Bluestein: n entries exp(-i pi k^2 / n)
    float2 *d_bf;        // Bluestein: L entries, spectrum of the wrapped conjugate chirp / L
};
struct B2sFftHostTables { std::vector<float2> tw, chirp, bf; };
// factorisation, batching and (when host != NULL) the tables the caller uploads into d_twiddle / d_chirp / d_bf
void b2s_fft_plan_init(B2sFftPlan *fp, int n, B2sFftHostTables *host);
// rows x n image; transform along the contiguous axis when along_cols == 0, along the row index otherwise.
// d_notch: n floats multiplying the packed (fftpack) spectrum; result overwrites the image.
void b2s_launch_notch(const B2sFftPlan &fp, const float *d_notch, const B2sImg &img, int along_cols, int n_planes,
                      int sm_count, cudaStream_t s);
size_t b2s_notch_smem(int n);

// rfft_exact.cu -------------------------------------------------------------------------------------------------
// rounding-exact mirror of scipy.fftpack.rfft / irfft (float32); nullptr when the length class is not covered
struct B2sXfftPlan;
int b2s_xfft_supported(int n);
B2sXfftPlan *b2s_xfft_create(int n);
void b2s_xfft_destroy(B2sXfftPlan *pl);
void b2s_launch_notch_exact(const B2sXfftPlan *pl, const float *d_notch, const B2sImg &img, int along_cols, int n_planes,
                            int sm_count, cudaStream_t s);

// bleach.cu ----------------------------------------------------------------------------------------------------
struct B2sBleachArgs {
    B2sImg img;                    // padded log-domain image; the rows x cols window starts at (base_pad, base_pad)
    int base_pad, rows, cols;
    double b0, b1, a1, zi;         // butter(1, f, output='sos') = [b0, b1, 0, 1, a1, 0]; sosfilt_zi(sos)[0, 0]
    double clip_min, clip_med, clip_max;
    const double *clip_pp;         // optional: (clip_min, clip_med, clip_max) per plane (multi-Otsu levels, core.py:1066-1077)
    double *scratch;               // forward-pass output, rows x (cols + 12) doubles per plane
    size_t scratch_plane_stride;   // doubles
    float *filt;                   // img_filter, rows x cols per plane
    unsigned *maxkey;              // per plane: order-preserving key of max(img_filter)
};
// correct_bleaching (core.py:501-559, non-max method) in place on the cropped window
void b2s_launch_bleach(const B2sBleachArgs &a, int n_planes, cudaStream_t s);
// max method (core.py:533-545); a.filt: 2 * (rows + cols) floats per plane, a.scratch: max(rows, cols) + 12 doubles per plane
void b2s_launch_bleach_max_method(const B2sBleachArgs &a, int n_planes, cudaStream_t s);

// lightsheet.cu -------------------------------------------------------------------------------------------------
struct B2sLightsheet;
const char *b2s_lightsheet_check(int rows, int cols, int artifact_length, int window);
B2sLightsheet *b2s_lightsheet_create(int rows, int cols, int dtype, int artifact_length, int window, double percentile,
                                     double weight, int weight_is_float);
void b2s_lightsheet_destroy(B2sLightsheet *L);
size_t b2s_lightsheet_grid_elems(const B2sLightsheet *L, int which /*0 ls grid, 1 bg grid, 2 sorted cell lists*/);
// mid: post-dark image in the lightsheet plan's dtype; e carries the final conversion / orientation / output
void b2s_launch_lightsheet(const B2sLightsheet *L, const void *mid, unsigned short *ls_grid, unsigned short *bg_grid,
                           unsigned short *cell_lists, const B2sEpilogueArgs &e, int n_planes, cudaStream_t s);
