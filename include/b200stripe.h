/*
 * b200stripe — C ABI of the B200-native (sm_100a) pystripe hot path.
 *
 * The reference (ucla-brain/image-preprocessing-pipeline) has no FFI layer for this path: its boundary is the
 * Python module surface of pystripe/core.py.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  Host bindings: image-preprocessing-pipeline_b200/pystripe/_native.py
 * (ctypes); the stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: every call returns 0 on success or a negative b2s_status; the message for the last failure on a
 * context is b2s_last_error(ctx).  Nothing throws across the ABI.  The caller owns every buffer it passes.
 * One context per GPU; a context and its plans may be used from one host thread at a time.
 * There is no CPU fallback: without a CUDA device b2s_create fails with B2S_ERR_CUDA.
 */
#ifndef B200STRIPE_H
#define B200STRIPE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100 /* 0.1.0 */

typedef enum b2s_status {
    B2S_OK = 0,
    B2S_ERR_INVALID = -1,     /* bad argument (ValueError / RuntimeError on the Python side)           */
    B2S_ERR_UNSUPPORTED = -2, /* valid in the reference, not implemented by this build                 */
    B2S_ERR_CUDA = -3,        /* CUDA runtime failure (message carries cudaGetErrorString)             */
    B2S_ERR_NOMEM = -4
} b2s_status;

typedef enum b2s_dtype { B2S_U8 = 0, B2S_U16 = 1, B2S_F32 = 2 } b2s_dtype;

/* numpy.pad modes accepted by filter_streaks (pystripe/core.py:1088-1089). */
typedef enum b2s_pad_mode {
    B2S_PAD_REFLECT = 0, B2S_PAD_WRAP = 1, B2S_PAD_SYMMETRIC = 2, B2S_PAD_EDGE = 3, B2S_PAD_CONSTANT = 4,
    /* modes whose pad area is computed from the image (numpy defaults: stat_length=None, end_values=0) */
    B2S_PAD_LINEAR_RAMP = 5, B2S_PAD_MAXIMUM = 6, B2S_PAD_MEAN = 7, B2S_PAD_MEDIAN = 8, B2S_PAD_MINIMUM = 9,
    B2S_PAD_EMPTY = 10 /* numpy leaves the pad area uninitialised; zeros here */
} b2s_pad_mode;

typedef enum b2s_ds_method { B2S_DS_MAX = 0, B2S_DS_MIN = 1, B2S_DS_MEAN = 2, B2S_DS_MEDIAN = 3 } b2s_ds_method;

/* debug_stop_after values (parity tests read intermediate coefficients with b2s_debug_read). */
typedef enum b2s_stage { B2S_STAGE_ALL = 0, B2S_STAGE_PROLOGUE = 1, B2S_STAGE_FORWARD = 2, B2S_STAGE_NOTCH = 3,
                         B2S_STAGE_INVERSE = 4 } b2s_stage;

/*
 * Plan parameters = the keyword arguments of process_img (pystripe/core.py:1190-1226) that reach arithmetic,
 * which include every argument of filter_streaks (pystripe/core.py:982-1002).
 * Zero-initialise, set struct_size = sizeof(b2s_params), then fill.
 */
typedef struct b2s_params {
    int32_t struct_size;
    int32_t height, width;        /* input plane (rows, cols)                                              */
    int32_t in_dtype;             /* b2s_dtype of the input planes (U16 or U8; F32 for filter_streaks(float)) */
    int32_t out_dtype;            /* b2s_dtype of the output planes                                        */
    /* --- filter_streaks ------------------------------------------------------------------------------- */
    double sigma1, sigma2;        /* sigma=(foreground, background); both 0 => destripe skipped (core.py:1058) */
    int32_t threshold_nonpositive;/* `threshold is not None and threshold <= 0`: one pass with sigma1 (core.py:946) */
    int32_t level;                /* 0 = maximum (pywt level=None), core.py:846                              */
    int32_t n_taps;               /* wavelet decomposition low-pass length F (even, 2..128)                  */
    const double *dec_lo;         /* host pointer, F doubles == pywt.Wavelet(name).dec_lo                    */
    int32_t pad_mode;             /* b2s_pad_mode                                                           */
    int32_t bidirectional;        /* core.py:1112-1117: also notch cV along axis -2                          */
    int32_t log1p;                /* log1p_normalization_needed (core.py:1063, 1150)                         */
    /* --- process_img ---------------------------------------------------------------------------------- */
    int32_t process_img;          /* 0: filter_streaks semantics only; 1: full process_img order of operations */
    int32_t has_flat;             /* flat-field supplied via b2s_plan_set_flat (core.py:1248-1250)           */
    int32_t gaussian;             /* gaussian_filter_2d (core.py:1280-1284, intended semantics)              */
    int32_t down_sample_y, down_sample_x; /* 0/1 = none (core.py:1286-1300); applied BEFORE the destripe     */
    int32_t down_sample_method;   /* b2s_ds_method                                                          */
    double dark;                  /* core.py:1324-1330                                                      */
    int32_t lightsheet;           /* core.py:1333-1348 -> pystripe/lightsheet_correct.py:31                  */
    int32_t artifact_length, background_window_size;
    double percentile, lightsheet_vs_background;
    int32_t convert_to_16bit, convert_to_8bit, bit_shift_to_right; /* core.py:1361-1369, 397-423              */
    int32_t rotate;               /* 0, 90, 180, 270 (core.py:1374-1379)                                     */
    int32_t flip_upside_down;     /* core.py:1371                                                           */
    int32_t reference_quirks;     /* 1: reproduce as-written behaviour (Gaussian result discarded)           */
    int32_t new_height, new_width;/* new_size (core.py:1356-1359): skimage.transform.resize(order 1) after the
                                     lightsheet stage, before the 8/16-bit conversion; 0 = none.  Pure up-sizing
                                     (anti_aliasing with sigma 0) and down-sizing (anti_aliasing=False) only     */
    /* --- bleach correction inside filter_streaks (core.py:501-559, 1131-1148) -------------------------- */
    int32_t bleach;               /* 1: correct_bleaching (non-max method) on the cropped log image, before expm1;
                                     2: max method (core.py:533-545): outer product of the low-passed row / column maxima:
                                     filter = sosfiltfilt(butter(1, frequency), clip(img with 0 -> clip_med)) along each
                                     row in float64 (scipy.signal, odd extension of 6 samples, sosfilt_zi start), cast to
                                     float32; img = img / filter * max(filter)                                    */
    int32_t bleach_per_plane;     /* 1: the three clip levels (and the constant-padding value) differ per plane and arrive through
                                     b2s_plan_set_bleach_levels before every b2s_run — the reference derives them per image with
                                     skimage.filters.threshold_multiotsu when they are not given (core.py:1066-1077) */
    double bleach_b0, bleach_b1, bleach_a1, bleach_zi; /* the one section butter(1, f, output='sos') returns:
                                     [b0, b1, 0, 1, a1, 0], and sosfilt_zi(sos)[0, 0] (host: scipy, core.py:495-497) */
    double bleach_clip_min, bleach_clip_med, bleach_clip_max; /* bounds as numpy.clip compares them (a weak Python
                                     float is rounded to float32 by the caller; log1p(1) stays float64, core.py:529-531) */
    double pad_constant;          /* padding_mode='constant': log1p(bleach_correction_clip_min) when that is given
                                     (core.py:1101-1105), else 0                                                  */
    int32_t aa_radius_y, aa_radius_x; /* new_size with anti_aliasing (skimage.transform.resize -> scipy.ndimage.
                                     gaussian_filter(sigma = (in/out - 1) / 2, mode='mirror', truncate 4) ahead of the zoom):
                                     kernel radius int(4 sigma + 0.5) per axis, 0 = no filter along that axis; the 2r+1
                                     weights are uploaded with b2s_plan_set_aa_weights                              */
    /* --- masking inside filter_streaks (core.py:475-489, 1079-1080) ------------------------------------ */
    int32_t mask;                 /* enable_masking with close_steps / open_steps given: img *= get_img_mask(img, threshold)
                                     on the (log) image ahead of the padding: img > threshold, cv2 MORPH_CLOSE with ones(close,
                                     close), MORPH_OPEN with ones(open, open), holes that no corner reaches filled (four
                                     4-connected cv2.floodFill)                                                    */
    int32_t mask_close, mask_open;/* close_steps, open_steps (>= 1)                                                 */
    int32_t mask_per_plane;       /* 1: one threshold per plane through b2s_plan_set_mask_thresholds before every b2s_run
                                     (bleach_correction_clip_med left to multi-Otsu, core.py:1066-1075)            */
    double mask_threshold;        /* bleach_correction_clip_med as numpy compares it with the image (the caller rounds a weak
                                     Python scalar to float32 against a float32 image); the kernel compares in double */
    /* --- execution ------------------------------------------------------------------------------------ */
    int32_t max_batch;            /* planes processed per launch group (workspace is sized for this)         */
    int32_t debug_stop_after;     /* b2s_stage; 0 in production                                             */
    int32_t exact;                /* 1 (default when 0 is passed through b2s_params_default): separate mul/add
                                     in the reference's summation order; 0 allows FMA contraction            */
} b2s_params;

typedef struct b2s_plan_info {
    int32_t out_height, out_width;      /* after down-sample / rotation                                     */
    int32_t out_dtype;                  /* b2s_dtype actually written (core.py:1361-1369 rules)             */
    int32_t n_passes;                   /* 0 (no destripe), 1, or 2 (sigma1 != sigma2, core.py:977-978)     */
    int32_t work_height, work_width;    /* image entering filter_streaks (after down-sample)                */
    int32_t base_pad, pad_y, pad_x;     /* core.py:1084-1096                                                */
    int32_t padded_height, padded_width;
    int32_t levels;
    int32_t level_rows[32], level_cols[32]; /* sub-band shape per level (index 0 = level 1)                 */
    int64_t workspace_bytes;
    int64_t algorithmic_bytes_per_plane;  /* SURVEY.md §8(d) stage model, for roofline reporting            */
    int64_t flops_per_plane;
} b2s_plan_info;

typedef struct b2s_context b2s_context;
typedef struct b2s_plan b2s_plan;

int b2s_version(void);
void b2s_params_default(b2s_params *p); /* defaults of filter_streaks: wrap padding, log1p, level 0, exact */

/* replaces: the implicit per-process device choice of the dead torch branch, pystripe/core.py:871-883, 1694-1695 */
int b2s_create(int device, b2s_context **ctx);
void b2s_destroy(b2s_context *ctx);
const char *b2s_last_error(const b2s_context *ctx);
int b2s_device_sm_count(const b2s_context *ctx);

/* replaces: argument handling of filter_streaks / process_img (core.py:1055-1110, 1227-1254) */
int b2s_plan_create(b2s_context *ctx, const b2s_params *params, b2s_plan **plan);
void b2s_plan_destroy(b2s_plan *plan);
int b2s_plan_query(const b2s_plan *plan, b2s_plan_info *info);
/* host-only: validate params and report the geometry a plan would have; needs no GPU (err may be NULL) */
int b2s_plan_geometry(const b2s_params *params, b2s_plan_info *info, char *err, size_t err_len);

/* GPU-free: the per-axis index / weight tables of the order-1 resize behind `new_size` (replaces the coordinate
 * set-up of scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True) that skimage.transform.resize reaches from
 * pystripe/core.py:1356-1359).  Output k reads source indices idx0[k], idx1[k] with float64 weights w0[k], w1[k]. */
int b2s_resize_table(int n_in, int n_out, int32_t *idx0, int32_t *idx1, double *w0, double *w1);
/* replaces: normalize_flat result captured in batch_filter's arg dict (core.py:1948-1953); flat is (height,width) f32 */
int b2s_plan_set_flat(b2s_plan *plan, const float *flat, int is_device);

/* replaces: np_notch (core.py:637-667) as evaluated by the host's numpy.  The plan builds its own float32 tables with
 * libm expf; numpy's float32 exp is a SIMD routine whose last bit differs from libm on ~40 % of arguments, so a host
 * that wants the reference's numpy values bit for bit uploads them here.  pass: 0 or 1 (sigma1 / sigma2 pass),
 * level: 1-based, axis: 0 = cH filtered along axis -1, 1 = cV along axis -2 (bidirectional); g: n host floats. */
/* replaces: scipy.ndimage._filters._gaussian_kernel1d as gaussian_filter calls it under skimage.transform.resize
 * (core.py:1356-1359).  axis 0 = rows (y), 1 = columns (x); n = 2 * aa_radius + 1 float64 weights, evaluated by the caller
 * with numpy exactly as scipy does (numpy's exp is not libm's). */
int b2s_plan_set_aa_weights(b2s_plan *plan, int axis, const double *weights, int n);
int b2s_plan_set_notch(b2s_plan *plan, int pass, int level, int axis, const float *g, int n);
/* replaces: np_filter_coefficient (core.py:749-754) for the one case the reference runs in float64 — integer pixels with
 * log1p_normalization_needed=False (pywt and scipy promote the integer image to double): irfft(rfft(x) * g) is applied as
 * the dense real n x n matrix it is.  R[k * n + c] = response at position c to a unit sample at position k, evaluated by the
 * caller with scipy.fftpack in float64 (host pointer).  b2s_plan_wants_notch_matrix tells whether a plan runs that way. */
int b2s_plan_wants_notch_matrix(const b2s_plan *plan);
int b2s_plan_set_notch_matrix(b2s_plan *plan, int pass, int level, int axis, const double *R, int n);
/* replaces: the per-image clip levels of filter_streaks when bleach_correction_clip_min / _med / _max are None
 * (core.py:1066-1077: lb, mb, ub = threshold_multiotsu(log1p(img), classes=4)).  clip: n_planes x 3 doubles (min, med, max as
 * numpy.clip compares them, after the clip_min >= log1p(1) rule of core.py:529-531); pad_value: n_planes floats,
 * log1p(clip_min) for padding_mode='constant' (core.py:1101-1105), or NULL.  Plane z of the next b2s_run uses entry z. */
int b2s_plan_set_bleach_levels(b2s_plan *plan, const double *clip, const float *pad_value, int64_t n_planes);
/* replaces: the threshold of get_img_mask when enable_masking is set and bleach_correction_clip_med is None
 * (core.py:1066-1075, 1079-1080: mb of threshold_multiotsu, per image).  thr: n_planes doubles; plane z of the next b2s_run
 * uses entry z.  Plans created with mask_per_plane only. */
int b2s_plan_set_mask_thresholds(b2s_plan *plan, const double *thr, int64_t n_planes);

/*
 * replaces: process_img(img, ...) / filter_streaks(img, ...) applied to n_planes independent planes
 * (core.py:1190, 982; driven per file by read_filter_save, core.py:1557).
 * in : n_planes x height x width, contiguous, params.in_dtype;  out: n_planes x out_height x out_width.
 * *_is_device = 1: pointer is device memory on the context's GPU (zero-copy torch tensors), work is enqueued on
 * `stream` (cudaStream_t, may be NULL) and the call returns without synchronising.
 * *_is_device = 0: host memory; the call stages through pinned buffers, overlaps H2D / compute / D2H on internal
 * streams and returns after the result is in `out`.
 */
int b2s_run(b2s_plan *plan, const void *in, void *out, int64_t n_planes, int in_is_device, int out_is_device,
            void *stream);

/* page-locked host allocations for the caller's staging buffers (host path runs at PCIe speed only from these) */
int b2s_host_alloc(b2s_context *ctx, size_t bytes, void **ptr);
int b2s_host_free(b2s_context *ctx, void *ptr);

/* number of kernels this library launched on the context since creation (bench.py reports it as gpu_launches) */
int64_t b2s_launch_count(const b2s_context *ctx);

/* event timing of the kernels inside b2s_run: when enabled every kernel launch is bracketed by CUDA events on the
 * launching stream; b2s_timing_read fills ms[class * B2S_TIMING_LEVELS + level] (accumulated milliseconds) and the
 * matching launch counts; level 0 collects launches that do not belong to a decomposition level. */
#define B2S_N_KERNEL_CLASSES 8
#define B2S_TIMING_LEVELS 33
enum { B2S_K_PRE = 0, B2S_K_PROLOGUE = 1, B2S_K_DWT_FWD = 2, B2S_K_NOTCH = 3, B2S_K_DWT_INV = 4, B2S_K_EPILOGUE = 5,
       B2S_K_LIGHTSHEET = 6, B2S_K_OTHER = 7 };
int b2s_timing_enable(b2s_context *ctx, int on);
int b2s_timing_read(b2s_context *ctx, double *ms_per_class, int64_t *launches_per_class, int reset);

/*
 * Parity-test hook: copy one float32 buffer of plane `plane` of the last batch to host.
 * what: 0 = padded image / reconstruction (padded_height x padded_width)
 *       1..4 = cA, cH, cV, cD of `level` (1-based)
 * out must hold rows*cols floats; rows/cols are returned.
 */
int b2s_debug_read(b2s_plan *plan, int what, int level, int plane, float *out, int32_t *rows, int32_t *cols);

/* device-side math used by the kernels, exposed so tests can compare them with the host libm bit for bit */
/* --- isotropic down-sampling of the post-stitch path (parallel_image_processor.py:371-435) -------------------------
 * b2s_isotropic_xy replaces, per plane: `img.astype(float32)`; for every (y_method, x_method) pair of
 * down_sampling_methods: block_reduce(img, (2, 1), y_method) while ceil(rows / 2) >= target_rows, block_reduce(img, (1, 2),
 * x_method) while ceil(cols / 2) >= target_cols (parallel_image_processor.py:376-381); then skimage.transform.resize(img,
 * target, preserve_range=True, anti_aliasing=True) (:383); uniform planes give zeros (:373-374).
 * steps: n_steps pairs (y_method, x_method), b2s_ds_method or -1 for None.  pre_rows / pre_cols: the shape the caller
 * expects ahead of the resize (checked).  wy / wx: the anti-aliasing Gaussian along each axis as numpy evaluates it
 * (2 * radius + 1 float64 weights, NULL / radius 0 = none).  d_in: n_planes planes of in_dtype, d_out: n_planes float32
 * planes of target_rows x target_cols, both DEVICE pointers; work is enqueued on `stream`. */
int b2s_isotropic_xy(b2s_context *ctx, const void *d_in, int in_dtype, int rows, int cols, int n_steps, const int32_t *steps,
                     int target_rows, int target_cols, int pre_rows, int pre_cols, const double *wy, int ry,
                     const double *wx, int rx, float *d_out, int n_planes, void *stream);
/* replaces: `resize(img, tile_size, preserve_range=True, anti_aliasing=True)` that read_filter_save applies to an input tile
 * whose shape differs from tile_size (pystripe/core.py:1540-1549): skimage.transform.resize = scipy.ndimage.gaussian_filter
 * (sigma = max(0, (in / out - 1) / 2) per axis, mode 'mirror', float64 for an integer image, float32 for a float32 one)
 * followed by the order-1 zoom and a clip to the image's own range.  wy / wx as for b2s_isotropic_xy.  d_in: n_planes device
 * planes of in_dtype; d_out: n_planes float32 device planes (the reference's float64 result, which filter_streaks casts to
 * float32 on entry). */
int b2s_resize_aa(b2s_context *ctx, const void *d_in, int in_dtype, int rows, int cols, int new_rows, int new_cols,
                  const double *wy, int ry, const double *wx, int rx, float *d_out, int n_planes, void *stream);
/* replaces: the z loop `block_reduce(z_stack, (2, 1, 1), z_method)` (parallel_image_processor.py:417-419) one level at
 * a time: d_out[k] = method(d_in[2k], d_in[2k+1]) over float32 planes of plane_elems elements; an odd last plane is paired
 * with zeros (cval = 0).  n_out = ceil(n_in / 2) planes are written. */
int b2s_isotropic_z(b2s_context *ctx, const float *d_in, int n_in, int64_t plane_elems, int method, float *d_out, void *stream);
/* replaces: the final conversion of the down-sampled plane (parallel_image_processor.py:422-433): mode 1 =
 * convert_to_16bit_fun (core.py:397-399), 2 = convert_to_8bit_fun with `shift` (core.py:402-423), 4 = astype(uint8).
 * d_out: uint16 (mode 1) or uint8. */
int b2s_isotropic_convert(b2s_context *ctx, const float *d_in, int64_t n, int mode, int shift, void *d_out, void *stream);

/* replaces: the deflate step of imsave_tif — tifffile.imwrite(..., compression=('ADOBE_DEFLATE', 1)), pystripe/core.py:275-334
 * (batch_filter's default output, core.py:1817) — on the device, so that D2H and the file write carry compressed bytes.
 * Every strip of rows_per_strip rows of every plane becomes one zlib stream (RFC 1950) holding one literal-only dynamic-Huffman
 * deflate block (RFC 1951); the streams are packed back to back into d_out in (plane, strip) order.  d_planes, d_out, d_sizes
 * (n_planes x strips uint32: bytes per stream) and d_offsets (n_planes x strips + 1 uint64: start of each stream, total last)
 * are DEVICE pointers; d_out must be 4-byte aligned with b2s_deflate_bound() bytes.  Synchronises the stream and returns the
 * packed size in *total_bytes.  b2sio_write_tiff_strips_batch (include/b2sio.h) lays such streams out as TIFF files. */
int64_t b2s_deflate_bound(int dtype, int rows, int cols, int n_planes, int rows_per_strip);
int b2s_deflate_strips(b2s_context *ctx, const void *d_planes, int dtype, int rows, int cols, int n_planes, int rows_per_strip,
                       void *d_out, int64_t out_capacity, uint32_t *d_sizes, uint64_t *d_offsets, int64_t *total_bytes, void *stream);
/* replaces: get_img_mask(img, threshold, close_steps, open_steps, flood_fill_flag=4) (pystripe/core.py:475-489) on n_planes
 * independent planes: img > threshold, cv2.morphologyEx MORPH_CLOSE with ones(close, close), MORPH_OPEN with ones(open, open),
 * background that no corner pixel reaches through 4-connected background added back (the four cv2.floodFill calls).
 * img, mask: DEVICE pointers (rows x cols per plane; mask bytes 0 / 1).  Synchronises the stream. */
int b2s_img_mask(b2s_context *ctx, const void *img, int dtype, int rows, int cols, int n_planes, double threshold, int close_steps,
                 int open_steps, unsigned char *mask, void *stream);
/* replaces: the pixel pass of `estimate_img_related_params` (process_images.py:594-659), which feeds log1p of a plane to
 * skimage.filters.threshold_multiotsu and to a masked percentile (:320-331).  Both are functions of the intensity
 * histogram: this entry ADDS the exact histogram of n_planes planes of uint8 / uint16 pixels (plane_elems each) into
 * `hist`: 65 536 uint64 counters when per_plane == 0, n_planes x 65 536 when per_plane != 0.  in / hist may be host or
 * device pointers.  Histograms are additive, so a whole stack sharded over GPUs needs one all-reduce of 65 536 counters
 * (the host side, pystripe/stack_stats.py, does it with torch.distributed). */
int b2s_histogram(b2s_context *ctx, const void *in, int in_is_device, int dtype, int64_t plane_elems, int n_planes,
                  uint64_t *hist, int hist_is_device, int per_plane, void *stream);

/* replaces: is_uniform_2d / is_uniform_3d (core.py:106-121) on a device array of n elements: *uniform = 1 when all equal */
int b2s_is_uniform(b2s_context *ctx, const void *d_in, int dtype, int64_t n, int32_t *uniform, void *stream);

int b2s_debug_math(b2s_context *ctx, int which /*0 log1pf, 1 expm1f*/, const float *in, float *out, int64_t n);
/* the fast epilogue maps a log value to its output integer — expm1f, [rint + clip of the integer path], dark, clip — through
 * 65 535 thresholds and an approximate exponential: compares that with the direct evaluation for the float bit patterns
 * [first, first + count) (all 2^32 in the GPU tests) for one combination of int_path / dark / work dtype */
int b2s_debug_expm1_table_check(b2s_context *ctx, int int_path, double dark, int work_dtype, uint64_t first, uint64_t count,
                                uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* B200STRIPE_H */
