// C ABI (include/b200stripe.h): contexts, plans (geometry, tables, workspace) and the batched run loop.
#include <emmintrin.h>
#include <sched.h>

#include <atomic>
#include <cctype>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200stripe.h"
#include "b2s_internal.h"

// Host staging pool: pageable caller buffers are copied to / from the slots' pinned staging buffers by a few worker
// threads (one memcpy thread moves ~10 GB/s, a PCIe 5 x16 link wants 2 x 46 GB/s), asynchronously to the run loop.
struct HostPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
    explicit HostPool(int n)
    {
        for (int i = 0; i < n; ++i)
            th.emplace_back([this] {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front());
                        q.pop_front();
                    }
                    f();
                }
            });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    void submit(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> lk(m); q.push_back(std::move(f)); }
        cv.notify_one();
    }
    int size() const { return (int)th.size(); }
};
struct TaskGroup {
    std::mutex m;
    std::condition_variable cv;
    int pending = 0;
    void add(int n) { std::lock_guard<std::mutex> lk(m); pending += n; }
    void done() { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return pending == 0; }); }
};

struct b2s_context {
    HostPool *pool = nullptr;
    int device = 0;
    int sm_count = 0;
    std::string err;
    int64_t launches = 0;
    bool timing = false;
    struct Span { int cls, level, n; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double ms[B2S_N_KERNEL_CLASSES][B2S_TIMING_LEVELS] = {{0}};
    int64_t n_launch[B2S_N_KERNEL_CLASSES][B2S_TIMING_LEVELS] = {{0}};
};

namespace {

int fail(b2s_context *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(ctx, call)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(ctx, B2S_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

int dtype_size(int d) { return d == B2S_U8 ? 1 : (d == B2S_U16 ? 2 : 4); }
int round4(int n) { return (n + 3) & ~3; }

// pystripe/core.py:670-698 ------------------------------------------------------------------------------------
double py_round2(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", v);  // correctly rounded decimal, like Python's round(v, 2)
    return strtod(buf, nullptr);
}
int notch_rise_point(double sigma, double rise)
{
    return (int)(std::sqrt(-2.0 * (sigma * sigma) * std::log(1.0 - rise)) + .5) / 2 * 2;
}
int calculate_pad_size(int rows, int cols, double sigma, double rise = 0.5)
{
    if (sigma == 0) return 0;
    const double x = cols + 1, y = rows + 1, c = 5e14;
    const double root = std::sqrt(x * x - 2 * x * y + y * y + 4 * c);
    const double r = py_round2(1 - std::exp((x + y - root) / (4 * (sigma * sigma)))) - 0.01;
    if (r < rise) rise = r;
    return notch_rise_point(sigma, rise);
}
int size_log2(size_t x) { int r = -1; while (x) { ++r; x >>= 1; } return r; }
int dwt_max_level(int n, int F) { return (F <= 1 || n < F - 1) ? 0 : size_log2((size_t)n / (F - 1)); }

struct Geometry {
    int in_rows, in_cols;
    int work_rows, work_cols;     // after down-sample
    int dsy, dsx;
    int base_pad, pad_y, pad_x, PH, PW;
    int n_passes;
    bool log_image;       // a log-domain working image exists: destripe passes and / or bleach correction
    double pass_sigma[2];
    int levels;
    int my[B2S_MAX_LEVELS + 1], mx[B2S_MAX_LEVELS + 1];  // index 0 = padded image, l = level l sub-bands
    int work_dtype;               // dtype of the array the reference holds when it enters filter_streaks
    int int_path;
    int final_mode, out_dtype;
    int out_rows, out_cols;
    bool fuse_flat;
    bool aa;                      // anti-aliasing Gaussian ahead of the zoom (aa_radius_y / aa_radius_x)
    bool resize;                  // new_size differs from the work image: order-1 zoom before the final conversion
    int new_rows, new_cols;
    int mid_dtype;                // dtype of the image the resize reads (after dark / lightsheet)
    bool f64;                     // integer pixels without log1p: the destripe runs in float64 (f64path.cu)
};

int compute_geometry(b2s_context *ctx, const b2s_params &p, Geometry &g)
{
    if (p.struct_size != (int)sizeof(b2s_params)) return fail(ctx, B2S_ERR_INVALID, "b2s_params.struct_size mismatch");
    if (p.height <= 0 || p.width <= 0) return fail(ctx, B2S_ERR_INVALID, "empty plane");
    if (p.in_dtype < B2S_U8 || p.in_dtype > B2S_F32) return fail(ctx, B2S_ERR_INVALID, "bad in_dtype");
    if (p.sigma1 < 0 || p.sigma2 < 0) return fail(ctx, B2S_ERR_INVALID, "np_notch: sigma must be positive");
    g.in_rows = p.height;
    g.in_cols = p.width;
    g.dsy = p.down_sample_y > 1 ? p.down_sample_y : 1;
    g.dsx = p.down_sample_x > 1 ? p.down_sample_x : 1;
    const bool ds = p.process_img && (p.down_sample_y > 0 || p.down_sample_x > 0) && (g.dsy > 1 || g.dsx > 1);
    g.work_rows = ds ? (p.height + g.dsy - 1) / g.dsy : p.height;
    g.work_cols = ds ? (p.width + g.dsx - 1) / g.dsx : p.width;
    if (p.process_img && ds && (p.down_sample_method < B2S_DS_MAX || p.down_sample_method > B2S_DS_MEDIAN))
        return fail(ctx, B2S_ERR_INVALID, "unsupported down-sampling method");
    if (p.process_img && ds && p.down_sample_method == B2S_DS_MEDIAN && g.dsy * g.dsx > 64)
        return fail(ctx, B2S_ERR_UNSUPPORTED, "median down-sampling over more than 64 samples per block is not implemented");

    // dtype bookkeeping (process_img order: flat -> gaussian -> block_reduce -> filter_streaks)
    int dt = p.in_dtype;
    const bool flat = p.process_img && p.has_flat;
    if (flat) dt = B2S_F32;
    const bool gauss = p.process_img && p.gaussian && !p.reference_quirks;
    if (gauss && dt == B2S_U8) return fail(ctx, B2S_ERR_UNSUPPORTED, "gaussian_filter_2d on uint8 is not implemented");
    if (ds && p.down_sample_method >= B2S_DS_MEAN) dt = B2S_F32;   // float64 in the reference; float32 holds what reaches log1p
    g.work_dtype = dt;

    // filter_streak_dual_band pass list, core.py:943-979
    const double s1 = p.sigma1, s2 = p.sigma2;
    if (s1 == 0 && s2 == 0) g.n_passes = 0;
    else if ((s1 > 0 && s1 == s2) || p.threshold_nonpositive) { g.n_passes = 1; g.pass_sigma[0] = s1; }
    else { g.n_passes = 2; g.pass_sigma[0] = s1; g.pass_sigma[1] = s2; }
    for (int i = 0; i < g.n_passes; ++i)
        if (g.pass_sigma[i] <= 0) return fail(ctx, B2S_ERR_INVALID, "np_notch: sigma must be positive");
    if (p.bleach) {   // correct_bleaching, core.py:501-559 (non-max method) — runs on the cropped reconstruction
        if (!p.log1p)
            return fail(ctx, B2S_ERR_UNSUPPORTED, "bleach correction needs log1p_normalization_needed=True (the reference's clip levels are log-domain)");
        if (g.work_cols <= 6 || (p.bleach == 2 && g.work_rows <= 6))
            return fail(ctx, B2S_ERR_INVALID, "The length of the input vector x must be greater than padlen, which is 6.");
    }
    g.log_image = g.n_passes > 0 || p.bleach;   // sigma = (0, 0) with a bleach frequency: log1p -> bleach -> expm1, no padding (core.py:1081, 1131)
    g.fuse_flat = flat && !gauss && !ds && g.log_image;   // the prologue divides by flat; without a destripe it is a pre-op

    g.base_pad = g.pad_y = g.pad_x = 0;
    g.PH = g.work_rows;
    g.PW = g.work_cols;
    g.levels = 0;
    g.my[0] = g.PH;
    g.mx[0] = g.PW;
    if (g.n_passes > 0) {
        if (p.pad_mode < B2S_PAD_REFLECT || p.pad_mode > B2S_PAD_EMPTY)
            return fail(ctx, B2S_ERR_INVALID, "Unsupported padding mode");
        if (!b2s_pad_fill_supported(p.pad_mode, g.work_rows, g.work_cols))
            return fail(ctx, B2S_ERR_UNSUPPORTED, "padding_mode='median' is limited to image sides of 32768 pixels");
        if (p.n_taps < 2 || p.n_taps > B2S_MAX_TAPS || (p.n_taps & 1) || !p.dec_lo)
            return fail(ctx, B2S_ERR_INVALID, "wavelet filter must have an even length in [2, %d]", B2S_MAX_TAPS);
        // core.py:1084-1096
        g.pad_y = g.work_rows % 2;
        g.pad_x = g.work_cols % 2;
        g.base_pad = calculate_pad_size(g.work_rows, g.work_cols, s1 > s2 ? s1 : s2);
        const int min_len = 34;
        if (g.work_rows + 2 * g.base_pad + g.pad_y < min_len) g.pad_y = min_len - (g.work_rows + 2 * g.base_pad);
        if (g.work_cols + 2 * g.base_pad + g.pad_x < min_len) g.pad_x = min_len - (g.work_cols + 2 * g.base_pad);
        g.PH = g.work_rows + 2 * g.base_pad + g.pad_y;
        g.PW = g.work_cols + 2 * g.base_pad + g.pad_x;
        const int F = p.n_taps;
        const int lmax = std::min(dwt_max_level(g.PH, F), dwt_max_level(g.PW, F));
        int L = p.level == 0 ? lmax : p.level;
        if (L < 0) return fail(ctx, B2S_ERR_INVALID, "Level value of %d is too low . Minimum level is 0.", L);
        if (L > B2S_MAX_LEVELS) return fail(ctx, B2S_ERR_INVALID, "level too high");
        g.my[0] = g.PH;
        g.mx[0] = g.PW;
        for (int l = 1; l <= L; ++l) {
            if (g.my[l - 1] < F || g.mx[l - 1] < F)
                return fail(ctx, B2S_ERR_UNSUPPORTED, "level %d: image side %dx%d shorter than the filter (%d taps)", l,
                            g.my[l - 1], g.mx[l - 1], F);
            g.my[l] = (g.my[l - 1] + F - 1) / 2;
            g.mx[l] = (g.mx[l - 1] + F - 1) / 2;
        }
        g.levels = L;
    }
    // new_size, core.py:1356-1359: `tile_size < new_size` / `>` compare (rows, cols) tuples lexicographically
    g.resize = false;
    g.aa = false;
    g.new_rows = g.work_rows;
    g.new_cols = g.work_cols;
    g.mid_dtype = g.work_dtype;
    if (p.process_img && p.new_height > 0 && p.new_width > 0 && (p.new_height != g.work_rows || p.new_width != g.work_cols)) {
        const bool up = g.work_rows < p.new_height || (g.work_rows == p.new_height && g.work_cols < p.new_width);
        // anti_aliasing=True (`tile_size < new_size`) filters with sigma = max(0, (in/out - 1) / 2) per axis: a new_size
        // larger along one axis and smaller along the other needs the caller's Gaussian (aa_radius_*, b2s_plan_set_aa_weights)
        if (up && ((g.work_rows > p.new_height && p.aa_radius_y <= 0) || (g.work_cols > p.new_width && p.aa_radius_x <= 0)))
            return fail(ctx, B2S_ERR_UNSUPPORTED,
                        "new_size larger along one axis and smaller along the other needs the anti-aliasing Gaussian of skimage.transform.resize (aa_radius_y / aa_radius_x)");
        if (p.aa_radius_y < 0 || p.aa_radius_x < 0 || p.aa_radius_y > 4096 || p.aa_radius_x > 4096)
            return fail(ctx, B2S_ERR_INVALID, "anti-aliasing radius out of range");
        if (g.work_dtype != B2S_F32 && p.dark > 0 && p.dark != std::floor(p.dark))
            return fail(ctx, B2S_ERR_UNSUPPORTED, "new_size after a fractional dark on an integer image (float64 in the reference) is not implemented");
        g.resize = true;
        g.aa = p.aa_radius_y > 0 || p.aa_radius_x > 0;
        g.new_rows = p.new_height;
        g.new_cols = p.new_width;
        if (p.lightsheet) g.mid_dtype = p.out_dtype;   // correct_lightsheet returns d_type
    }
    g.int_path = g.log_image && g.work_dtype != B2S_F32;
    g.f64 = g.int_path && !p.log1p;   // pywt / scipy promote the integer image to float64 (core.py:1063, 1081-1158)
    if (g.f64 && (p.bleach || g.n_passes == 0 || p.pad_mode > B2S_PAD_CONSTANT))
        return fail(ctx, B2S_ERR_UNSUPPORTED,
                    "log1p_normalization_needed=False on an integer image: only the destripe with a copying padding mode is implemented");
    if (p.mask && (p.mask_close < 1 || p.mask_open < 1))
        return fail(ctx, B2S_ERR_INVALID, "enable_masking: close_steps and open_steps must be at least 1 (cv2 rejects an empty structuring element)");

    // final conversion, core.py:1361-1369
    if (!p.process_img) {
        g.out_dtype = p.in_dtype;
        g.final_mode = p.in_dtype == B2S_F32 ? 3 : 0;
    } else {
        // dtype of the array right before the conversion
        int cur = g.work_dtype;
        if (p.dark > 0 && cur != B2S_F32 && p.dark != std::floor(p.dark)) cur = B2S_F32;  // promoted to float64
        if (g.resize) cur = B2S_F32;   // skimage.transform.resize returns float64 (float32 for a float32 image)
        if (p.convert_to_16bit && cur != B2S_U16) { g.final_mode = 1; g.out_dtype = B2S_U16; }
        else if (p.convert_to_8bit && cur != B2S_U8) { g.final_mode = 2; g.out_dtype = B2S_U8; }
        else if (p.out_dtype != B2S_F32) { g.final_mode = 0; g.out_dtype = p.out_dtype; }
        else { g.final_mode = 3; g.out_dtype = B2S_F32; }
        if (p.convert_to_8bit && (p.bit_shift_to_right < 0 || p.bit_shift_to_right > 8))
            return fail(ctx, B2S_ERR_INVALID, "right shift should be between 0 and 8");
    }
    if (p.rotate != 0 && p.rotate != 90 && p.rotate != 180 && p.rotate != 270)
        return fail(ctx, B2S_ERR_INVALID, "rotate must be 0, 90, 180 or 270");
    const bool swap = p.process_img && (p.rotate == 90 || p.rotate == 270);
    g.out_rows = swap ? g.new_cols : g.new_rows;
    g.out_cols = swap ? g.new_rows : g.new_cols;
    if (p.process_img && p.lightsheet) {
        if (const char *why = b2s_lightsheet_check(g.work_rows, g.work_cols, p.artifact_length, p.background_window_size))
            return fail(ctx, B2S_ERR_UNSUPPORTED, "%s", why);
        if (!(p.out_dtype == B2S_U16 || (p.out_dtype == B2S_U8 && g.work_dtype == B2S_U8)))
            return fail(ctx, B2S_ERR_UNSUPPORTED, "lightsheet: d_type must be uint16 (or uint8 for uint8 input)");
        if (g.work_dtype != B2S_F32 && p.dark > 0 && p.dark != std::floor(p.dark))
            return fail(ctx, B2S_ERR_UNSUPPORTED, "lightsheet after a fractional dark on an integer image (float64 in the reference) is not implemented");
        if (p.percentile < 0 || p.percentile > 1) return fail(ctx, B2S_ERR_INVALID, "percentile must be in [0, 1]");
    }
    return B2S_OK;
}

void fill_info(const b2s_params &p, const Geometry &g, b2s_plan_info *info)
{
    memset(info, 0, sizeof *info);
    info->out_height = g.out_rows;
    info->out_width = g.out_cols;
    info->out_dtype = g.out_dtype;
    info->n_passes = g.n_passes;
    info->work_height = g.work_rows;
    info->work_width = g.work_cols;
    info->base_pad = g.base_pad;
    info->pad_y = g.pad_y;
    info->pad_x = g.pad_x;
    info->padded_height = g.PH;
    info->padded_width = g.PW;
    info->levels = g.levels;
    int64_t bytes = 0, flops = 0;
    const int F = p.n_taps;
    if (g.n_passes > 0) {
        bytes += (int64_t)dtype_size(g.work_dtype) * g.work_rows * g.work_cols + 4ll * g.PH * g.PW;  // prologue
        int64_t per_pass = 0;
        for (int l = 1; l <= g.levels; ++l) {
            info->level_rows[l - 1] = g.my[l];
            info->level_cols[l - 1] = g.mx[l];
            const int64_t sub = (int64_t)g.my[l] * g.mx[l], inp = (int64_t)g.my[l - 1] * g.mx[l - 1];
            per_pass += 4 * inp + 16 * sub;                       // forward
            per_pass += 8 * sub * (p.bidirectional ? 2 : 1);      // notch
            per_pass += 16 * sub + 4 * inp;                       // inverse
            // MACs: axis -2 pass 2F per (my x nx) sample, axis -1 pass 4F per (my x mx); synthesis mirrors it
            flops += 2 * 2 * ((int64_t)2 * F * g.my[l] * g.mx[l - 1] + (int64_t)4 * F * sub);
        }
        bytes += per_pass * g.n_passes;
        flops *= g.n_passes;
        bytes += 4ll * g.work_rows * g.work_cols + (int64_t)dtype_size(g.out_dtype) * g.out_rows * g.out_cols;  // epilogue
    } else {
        bytes += (int64_t)(dtype_size(p.in_dtype)) * p.height * p.width + (int64_t)dtype_size(g.out_dtype) * g.out_rows * g.out_cols;
    }
    info->algorithmic_bytes_per_plane = bytes;
    info->flops_per_plane = flops;
}

}  // namespace

struct b2s_plan {
    b2s_context *ctx = nullptr;
    b2s_params p;
    Geometry g;
    B2sTaps taps;
    int B = 1;            // max batch
    // device workspace (one slot per in-flight batch)
    static constexpr int kSlots = 5;
    struct Slot {
        float *padded = nullptr;
        float *sub[B2S_MAX_LEVELS + 1][4] = {};
        void *d_in = nullptr, *d_out = nullptr;
        void *pre_a = nullptr, *pre_b = nullptr;   // pre-op temporaries
        float *dwt_scratch = nullptr;              // long filters: per-axis intermediates (b2s_dwt_scratch_floats per plane)
        void *mid = nullptr;                       // lightsheet / resize: post-dark image
        void *mid2 = nullptr;                      // lightsheet followed by resize: the cleaned image
        unsigned *mm2 = nullptr;                   // resize: per-plane min / max keys of the image it reads
        void *aa_a = nullptr, *aa_b = nullptr;     // resize with anti-aliasing: the filtered image after each axis (f64 / f32)
        unsigned short *ls_grid = nullptr, *bg_grid = nullptr, *ls_cells = nullptr;
        unsigned *mm = nullptr;
        unsigned *pad_flags = nullptr;             // computed padding modes: 4 flags per plane
        double *f64_work = nullptr;                // f64 path: padded image, sub-bands, intermediates (doubles per plane x B)
        double *bleach_scratch = nullptr;          // bleach correction: forward low-pass output, rows x (cols + 12) per plane
        float *bleach_filt = nullptr;              //                    img_filter, rows x cols per plane
        unsigned *bleach_max = nullptr;            //                    per-plane key of max(img_filter)
        unsigned char *mask_a = nullptr, *mask_b = nullptr, *mask_c = nullptr;   // enable_masking: mask, scratch, flood reach (bytes per pixel)
        int *mask_flag = nullptr, *mask_hflag = nullptr;                         //                 flood-fill "changed" flag (device / page-locked)
        void *h_in = nullptr, *h_out = nullptr;    // pinned staging
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
    } slot[kSlots];
    int n_slots = kSlots;            // slots in use: fewer when kSlots x B planes of workspace do not fit the device
    bool dry = false;                // sizing pass: dev_alloc only adds up
    size_t dry_bytes = 0;
    size_t dwt_scratch_stride = 0;   // floats per plane of Slot::dwt_scratch (0: fused DWT kernels)
    cudaEvent_t fork = nullptr;      // device-resident runs: the caller's stream position the slot streams wait for
    int pitch[B2S_MAX_LEVELS + 1];
    size_t plane_stride[B2S_MAX_LEVELS + 1];
    float *d_flat = nullptr;
    // numpy.pad tables for the prologue
    int n_row_groups = 0;
    int *d_row_src = nullptr, *d_row_start = nullptr, *d_row_targets = nullptr, *d_colmap = nullptr;
    float *d_lut = nullptr;
    float *d_epi_thr = nullptr;     // fast epilogue: thresholds of the plan's log value -> integer map (pointwise.cu)
    int epi_kmax = 0;
    double taps64[4][B2S_MAX_TAPS] = {};            // f64 path: dec_lo, dec_hi, rec_lo, rec_hi in double
    double *d_notch_mat[2][B2S_MAX_LEVELS + 1][2] = {};   // f64 path: n x n response matrices (b2s_plan_set_notch_matrix)
    size_t f64_plane_doubles = 0;
    double *d_clip_pp = nullptr;                   // bleach_per_plane: (min, med, max) per plane of the coming b2s_run
    float *d_padv_pp = nullptr;                     //                   constant-padding value per plane (or null)
    int64_t n_levels = 0, cap_levels = 0;
    double *d_mask_thr_pp = nullptr;   // enable_masking with per-plane thresholds (multi-Otsu clip_med)
    int64_t n_mask_thr = 0, cap_mask_thr = 0;
    int *d_rz_idx = nullptr;                        // new_size: [iy0 | iy1 | ix0 | ix1]
    double *d_rz_w = nullptr;                       //           [wy0 | wy1 | wx0 | wx1]
    double *d_aa_w[2] = {nullptr, nullptr};         // anti-aliasing Gaussian weights per axis (2 r + 1), caller-supplied
    std::map<int, B2sFftPlan> fft;                  // by length
    B2sLightsheet *ls = nullptr;
    std::map<int, B2sXfftPlan *> xfft;              // by length: rounding-exact transform (exact mode, covered lengths)
    float *d_notch[2][B2S_MAX_LEVELS + 1][2] = {};  // [pass][level][axis: 0 = cH rows, 1 = cV cols]
    int64_t workspace_bytes = 0;
    std::vector<void *> allocs;
    int last_batch = 0;
};

namespace {

int dev_alloc(b2s_plan *pl, void **ptr, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (pl->dry) { pl->dry_bytes += (bytes + 511) & ~(size_t)511; return B2S_OK; }
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return fail(pl->ctx, B2S_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    pl->allocs.push_back(*ptr);
    pl->workspace_bytes += (int64_t)bytes;
    return B2S_OK;
}

// brackets `n_launches` kernel launches of class `k` (level 0 = not level-specific) with CUDA events on the
// launching stream when timing is enabled; always counts the launches
struct ClassTimer {
    b2s_context *ctx;
    cudaStream_t s;
    int cls, level, n;
    cudaEvent_t a = nullptr;
    ClassTimer(b2s_context *c, cudaStream_t st, int k, int n_launches, int lvl = 0)
        : ctx(c), s(st), cls(k), level(lvl), n(n_launches)
    {
        ctx->launches += n;
        if (ctx->timing) { cudaEventCreate(&a); cudaEventRecord(a, s); }
    }
    ~ClassTimer()
    {
        if (ctx->timing) {
            cudaEvent_t b;
            cudaEventCreate(&b);
            cudaEventRecord(b, s);
            ctx->spans.push_back({cls, level, n, a, b});
        }
    }
};

B2sImg img_of(b2s_plan *pl, float *base, int level)
{
    B2sImg im;
    im.ptr = base;
    im.plane_stride = pl->plane_stride[level];
    im.pitch = pl->pitch[level];
    im.rows = pl->g.my[level];
    im.cols = pl->g.mx[level];
    return im;
}

int build_resize_tables(b2s_plan *pl)
{
    b2s_context *ctx = pl->ctx;
    const Geometry &g = pl->g;
    if (g.resize) {   // axis tables of the order-1 zoom (scipy.ndimage NI_ZoomShift, grid_mode, mirror)
        const int nr = g.new_rows, nc = g.new_cols;
        std::vector<int> idx(2 * (size_t)(nr + nc));
        std::vector<double> w(2 * (size_t)(nr + nc));
        b2s_resize_axis_table(g.work_rows, nr, idx.data(), idx.data() + nr, w.data(), w.data() + nr);
        b2s_resize_axis_table(g.work_cols, nc, idx.data() + 2 * nr, idx.data() + 2 * nr + nc, w.data() + 2 * nr, w.data() + 2 * nr + nc);
        int rc = dev_alloc(pl, (void **)&pl->d_rz_idx, sizeof(int) * idx.size());
        if (rc) return rc;
        if ((rc = dev_alloc(pl, (void **)&pl->d_rz_w, sizeof(double) * w.size()))) return rc;
        CU(ctx, cudaMemcpy(pl->d_rz_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(pl->d_rz_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice));
    }
    return B2S_OK;
}

int build_tables(b2s_plan *pl)
{
    b2s_context *ctx = pl->ctx;
    const Geometry &g = pl->g;
    const b2s_params &p = pl->p;
    // filters: pywt derives everything from one table (wavelets.c); float32 copies of the double values
    const int F = p.n_taps;
    B2sTaps &t = pl->taps;
    memset(&t, 0, sizeof t);
    t.F = F;
    std::vector<double> dec_lo(F), rec_lo(F), rec_hi(F), dec_hi(F);
    for (int k = 0; k < F; ++k) dec_lo[k] = p.dec_lo[k];
    for (int k = 0; k < F; ++k) rec_lo[k] = dec_lo[F - 1 - k];
    for (int k = 0; k < F; ++k) rec_hi[k] = ((k & 1) ? -1.0 : 1.0) * rec_lo[F - 1 - k];
    for (int k = 0; k < F; ++k) dec_hi[k] = rec_hi[F - 1 - k];
    for (int k = 0; k < F; ++k) {
        t.dec_lo[k] = (float)dec_lo[k];
        t.dec_hi[k] = (float)dec_hi[k];
        t.rec_lo[k] = (float)rec_lo[k];
        t.rec_hi[k] = (float)rec_hi[k];
        pl->taps64[0][k] = dec_lo[k]; pl->taps64[1][k] = dec_hi[k]; pl->taps64[2][k] = rec_lo[k]; pl->taps64[3][k] = rec_hi[k];
    }
    // numpy.pad as index tables (core.py:1100-1110)
    {
        auto pad_index = [&](int i, int n) -> int {
            if (i >= 0 && i < n) return i;
            auto imod = [](int v, int m) { int t = v % m; return t < 0 ? t + m : t; };
            switch (p.pad_mode) {
            case B2S_PAD_REFLECT: { if (n == 1) return 0; const int q = 2 * (n - 1), t = imod(i, q); return t < n ? t : q - t; }
            case B2S_PAD_SYMMETRIC: { const int q = 2 * n, t = imod(i, q); return t < n ? t : q - 1 - t; }
            case B2S_PAD_WRAP: return imod(i, n);
            case B2S_PAD_EDGE: return i < 0 ? 0 : n - 1;
            default: return -1;
            }
        };
        std::vector<int> colmap(pl->pitch[0], -1);
        for (int c = 0; c < g.PW; ++c) colmap[c] = pad_index(c - g.base_pad, g.work_cols);
        std::vector<std::vector<int>> targets(g.work_rows);
        std::vector<int> orphan;
        for (int i = 0; i < g.PH; ++i) {
            const int sy = pad_index(i - g.base_pad, g.work_rows);
            if (sy >= 0) targets[sy].push_back(i); else orphan.push_back(i);
        }
        std::vector<int> row_src, row_start(1, 0), row_targets;
        for (int y = 0; y < g.work_rows; ++y) {
            if (targets[y].empty()) continue;
            row_src.push_back(y);
            row_targets.insert(row_targets.end(), targets[y].begin(), targets[y].end());
            row_start.push_back((int)row_targets.size());
        }
        for (int i : orphan) { row_src.push_back(-1); row_targets.push_back(i); row_start.push_back((int)row_targets.size()); }
        pl->n_row_groups = (int)row_src.size();
        struct { int **dst; std::vector<int> *src; } up[4] = {{&pl->d_row_src, &row_src}, {&pl->d_row_start, &row_start},
                                                              {&pl->d_row_targets, &row_targets}, {&pl->d_colmap, &colmap}};
        for (auto &u : up) {
            int rc = dev_alloc(pl, (void **)u.dst, sizeof(int) * u.src->size());
            if (rc) return rc;
            CU(ctx, cudaMemcpy(*u.dst, u.src->data(), sizeof(int) * u.src->size(), cudaMemcpyHostToDevice));
        }
        if (g.work_dtype != B2S_F32 && p.log1p) {   // integer pixels enter filter_streaks: table look-up
            int rc = dev_alloc(pl, (void **)&pl->d_lut, sizeof(float) * 65536);
            if (rc) return rc;
            b2s_launch_log1p_lut(pl->d_lut, 65536, 0);
            CU(ctx, cudaDeviceSynchronize());
        }
    }
    {
        static const bool no_tab = getenv("B2S_EPILOGUE_TABLE") && atoi(getenv("B2S_EPILOGUE_TABLE")) == 0;
        const double dark = p.process_img ? p.dark : 0.0;
        const bool f32_exact = !(dark > 0.0 && g.int_path && dark != std::floor(dark)) && dark < 16777216.0;
        const bool rotated = p.process_img && p.rotate != 0;
        if (g.log_image && p.log1p && !g.f64 && f32_exact && !rotated && g.final_mode != 3 && !no_tab && !pl->dry) {
            int rc = dev_alloc(pl, (void **)&pl->d_epi_thr, sizeof(float) * 65536 + sizeof(int));
            if (rc) return rc;
            B2sEpiFn f;
            f.int_path = g.int_path; f.darkf = (float)dark; f.hi_w = g.work_dtype == B2S_U8 ? 255.f : 65535.f;
            int *d_kmax = reinterpret_cast<int *>(pl->d_epi_thr + 65536);
            b2s_launch_epi_thresholds(pl->d_epi_thr, f, d_kmax, 0);
            CU(ctx, cudaMemcpy(&pl->epi_kmax, d_kmax, sizeof(int), cudaMemcpyDeviceToHost));
        }
    }
    // per-level FFT plans + notch tables (np_notch, core.py:637-667, numpy float32 branch); the float64 path applies the
    // notch as a dense matrix the caller uploads (b2s_plan_set_notch_matrix)
    for (int pass = 0; pass < (g.f64 ? 0 : g.n_passes); ++pass) {
        for (int l = 1; l <= g.levels; ++l) {
            for (int axis = 0; axis < (p.bidirectional ? 2 : 1); ++axis) {
                const int n = axis == 0 ? g.mx[l] : g.my[l];
                const int other = axis == 0 ? g.my[l] : g.mx[l];
                const int img_len = axis == 0 ? g.PH : g.PW;
                static const bool no_xfft = getenv("B2S_NO_XFFT") != nullptr;
                if (!pl->fft.count(n) && !pl->xfft.count(n) && p.exact && !no_xfft) {
                    // exact mode: the scipy.fftpack mirror serves every covered length (whole stitched slices: 5000+
                    // points); the FMA transform of fft.cu is only built for lengths the mirror does not cover
                    B2sXfftPlan *xp = b2s_xfft_create(n);
                    if (xp) pl->xfft[n] = xp;
                    CU(ctx, cudaGetLastError());
                }
                if (!pl->fft.count(n) && !pl->xfft.count(n)) {
                    if (b2s_notch_smem(n) > 220 * 1024)
                        return fail(ctx, B2S_ERR_UNSUPPORTED, "sub-band side %d too long for the shared-memory FFT", n);
                    B2sFftPlan fp;
                    B2sFftHostTables ht;
                    b2s_fft_plan_init(&fp, n, &ht);
                    fp.d_twiddle = fp.d_chirp = fp.d_bf = nullptr;
                    struct { float2 **dst; std::vector<float2> *src; } up[3] = {{&fp.d_twiddle, &ht.tw}, {&fp.d_chirp, &ht.chirp}, {&fp.d_bf, &ht.bf}};
                    for (auto &u : up) {
                        if (u.src->empty()) continue;
                        int rc = dev_alloc(pl, (void **)u.dst, sizeof(float2) * u.src->size());
                        if (rc) return rc;
                        CU(ctx, cudaMemcpy(*u.dst, u.src->data(), sizeof(float2) * u.src->size(), cudaMemcpyHostToDevice));
                    }
                    pl->fft[n] = fp;
                }
                const double width_frac = g.pass_sigma[pass] / (double)img_len;
                const double sigma_l = (double)other * width_frac;
                const float denom = -2.0f * (float)(sigma_l * sigma_l);
                std::vector<float> gt(n);
                for (int k = 0; k < n; ++k) {
                    float v = (float)k;
                    v = v * v;
                    v = v / denom;
                    v = expf(v);
                    gt[k] = 1.0f - v;
                }
                int rc = dev_alloc(pl, (void **)&pl->d_notch[pass][l][axis], sizeof(float) * n);
                if (rc) return rc;
                CU(ctx, cudaMemcpy(pl->d_notch[pass][l][axis], gt.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
            }
        }
    }
    return B2S_OK;
}

int alloc_slot(b2s_plan *pl, int si)
{
    b2s_context *ctx = pl->ctx;
    const Geometry &g = pl->g;
    const b2s_params &p = pl->p;
    b2s_plan::Slot &s = pl->slot[si];
    const size_t B = pl->B;
    int rc;
    if (g.f64) {
        pl->f64_plane_doubles = b2s_f64_workspace_doubles(g.PH, g.PW, g.levels, g.my, g.mx);
        if ((rc = dev_alloc(pl, (void **)&s.f64_work, sizeof(double) * pl->f64_plane_doubles * B))) return rc;
    }
    if (g.log_image) {
        if ((rc = dev_alloc(pl, (void **)&s.padded, sizeof(float) * pl->plane_stride[0] * B))) return rc;
        for (int l = 1; l <= (g.f64 ? 0 : g.levels); ++l)
            for (int k = 0; k < 4; ++k)
                if ((rc = dev_alloc(pl, (void **)&s.sub[l][k], sizeof(float) * pl->plane_stride[l] * B))) return rc;
        if (g.n_passes > 0 && p.pad_mode >= B2S_PAD_LINEAR_RAMP &&
            (rc = dev_alloc(pl, (void **)&s.pad_flags, sizeof(unsigned) * 4 * B))) return rc;
        pl->dwt_scratch_stride = (g.n_passes > 0 && !g.f64) ? b2s_dwt_scratch_floats(pl->taps.F, g.PH, g.PW) : 0;
        if (pl->dwt_scratch_stride &&
            (rc = dev_alloc(pl, (void **)&s.dwt_scratch, sizeof(float) * pl->dwt_scratch_stride * B))) return rc;
    }
    const size_t in_elems = (size_t)g.in_rows * g.in_cols, work_elems = (size_t)g.work_rows * g.work_cols;
    if (p.mask && g.log_image) {
        for (unsigned char **m : {&s.mask_a, &s.mask_b, &s.mask_c})
            if ((rc = dev_alloc(pl, (void **)m, work_elems * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.mask_flag, sizeof(int)))) return rc;
        if (!pl->dry && !s.mask_hflag) CU(pl->ctx, cudaMallocHost((void **)&s.mask_hflag, sizeof(int)));
    }
    if (p.bleach) {
        if ((rc = dev_alloc(pl, (void **)&s.bleach_scratch, sizeof(double) * (size_t)g.work_rows * (g.work_cols + 12) * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.bleach_filt, sizeof(float) * work_elems * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.bleach_max, sizeof(unsigned) * B))) return rc;
    }
    if ((rc = dev_alloc(pl, &s.d_in, in_elems * dtype_size(p.in_dtype) * B))) return rc;
    if ((rc = dev_alloc(pl, &s.d_out, (size_t)g.out_rows * g.out_cols * dtype_size(g.out_dtype) * B))) return rc;
    if (p.process_img) {
        if ((rc = dev_alloc(pl, &s.pre_a, in_elems * 4 * B))) return rc;
        if ((rc = dev_alloc(pl, &s.pre_b, in_elems * 4 * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.mm, sizeof(unsigned) * 2 * B))) return rc;
    }
    if (g.resize) {
        if (!pl->ls && (rc = dev_alloc(pl, &s.mid, work_elems * 4 * B))) return rc;
        if (pl->ls && (rc = dev_alloc(pl, &s.mid2, work_elems * 4 * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.mm2, sizeof(unsigned) * 2 * B))) return rc;
        if (g.aa) {   // skimage filters a float64 copy of an integer image, a float32 image stays float32
            const size_t esz = g.mid_dtype == B2S_F32 ? 4 : 8;
            if ((rc = dev_alloc(pl, &s.aa_a, work_elems * esz * B))) return rc;
            if (p.aa_radius_y > 0 && p.aa_radius_x > 0 && (rc = dev_alloc(pl, &s.aa_b, work_elems * esz * B))) return rc;
        }
    }
    if (pl->ls) {
        if ((rc = dev_alloc(pl, &s.mid, work_elems * 4 * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.ls_grid, sizeof(unsigned short) * b2s_lightsheet_grid_elems(pl->ls, 0) * B))) return rc;
        if ((rc = dev_alloc(pl, (void **)&s.bg_grid, sizeof(unsigned short) * b2s_lightsheet_grid_elems(pl->ls, 1) * B))) return rc;
        if (b2s_lightsheet_grid_elems(pl->ls, 2) &&
            (rc = dev_alloc(pl, (void **)&s.ls_cells, sizeof(unsigned short) * b2s_lightsheet_grid_elems(pl->ls, 2) * B))) return rc;
    }
    if (pl->dry) return B2S_OK;
    CU(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    return B2S_OK;
}

// Whole stitched slices (20k x 15k planes, ~7 GB of workspace each) do not fit kSlots x max_batch times:
// the batch is halved, then slots are dropped, until the workspace fits B2S_WORKSPACE_FRACTION (default 0.7) of the
// free device memory.  The caller's max_batch is a ceiling, b2s_run loops over whatever batch the plan ended up with.
int fit_workspace(b2s_plan *pl)
{
    b2s_context *ctx = pl->ctx;
    size_t free_b = 0, total_b = 0;
    CU(ctx, cudaMemGetInfo(&free_b, &total_b));
    const char *fe = getenv("B2S_WORKSPACE_FRACTION");
    double frac = fe ? atof(fe) : 0.7;
    if (!(frac > 0.0 && frac <= 1.0)) frac = 0.7;
    const double budget = frac * (double)free_b;
    auto slot_bytes = [&](int B) {
        const int keep = pl->B;
        pl->B = B; pl->dry = true; pl->dry_bytes = 0;
        alloc_slot(pl, 0);
        pl->slot[0] = b2s_plan::Slot();
        pl->dry = false; pl->B = keep;
        return (double)pl->dry_bytes;
    };
    int B = pl->B;
    while (B > 1 && b2s_plan::kSlots * slot_bytes(B) > budget) B = (B + 1) / 2;
    int ns = b2s_plan::kSlots;
    const double per_slot = slot_bytes(B);
    while (ns > 1 && ns * per_slot > budget) --ns;
    if (ns * per_slot > budget)
        return fail(ctx, B2S_ERR_NOMEM, "one plane needs %.1f GB of workspace, %.1f GB of device memory are free",
                    per_slot / 1e9, (double)free_b / 1e9);
    pl->B = B;
    pl->n_slots = ns;
    return B2S_OK;
}

// enqueue the whole per-batch pipeline on `st` using slot workspace `s`
int enqueue_batch(b2s_plan *pl, b2s_plan::Slot &s, const void *d_in, void *d_out, int nb, cudaStream_t st, int64_t z0 = 0)
{
    b2s_context *ctx = pl->ctx;
    const Geometry &g = pl->g;
    const b2s_params &p = pl->p;
    const int exact = p.exact;
    pl->last_batch = nb;

    // ---- process_img pre-ops (core.py:1232-1300)
    const void *cur = d_in;
    int cur_dt = p.in_dtype;
    // uniform-plane check (core.py:1232): min / max keys of the raw input.  When the prologue reads the raw input itself
    // (no pre-op in between) it accumulates them on the way; otherwise one vectorised pass does.
    const bool fuse_minmax = p.process_img && g.log_image && !g.f64 && !(p.gaussian && !p.reference_quirks) &&
                             g.work_rows == g.in_rows && g.work_cols == g.in_cols && (!(p.process_img && p.has_flat) || g.fuse_flat);
    if (p.process_img) {
        CU(ctx, cudaMemsetAsync(s.mm, 0xff, sizeof(unsigned) * 2 * nb, st));
        if (!fuse_minmax) {
            ClassTimer t(ctx, st, B2S_K_PRE, 1);
            b2s_launch_minmax(d_in, p.in_dtype, (size_t)g.in_rows * g.in_cols, nb, s.mm, st);
        }
    }
    const bool flat = p.process_img && p.has_flat;
    if (flat && !pl->d_flat) return fail(ctx, B2S_ERR_INVALID, "has_flat is set but b2s_plan_set_flat was not called");
    if (flat && !g.fuse_flat) {
        ClassTimer t(ctx, st, B2S_K_PRE, 1);
        b2s_launch_flat_divide(cur, cur_dt, pl->d_flat, (float *)s.pre_a, (size_t)g.in_rows * g.in_cols, nb, st);
        cur = s.pre_a;
        cur_dt = B2S_F32;
    }
    if (p.process_img && p.gaussian && !p.reference_quirks) {
        ClassTimer t(ctx, st, B2S_K_PRE, 1);
        if (cur_dt == B2S_F32) b2s_launch_gauss5_f32((const float *)cur, (float *)s.pre_b, g.in_rows, g.in_cols, nb, st);
        else b2s_launch_gauss5_u16((const uint16_t *)cur, (uint16_t *)s.pre_b, g.in_rows, g.in_cols, nb, st);
        cur = s.pre_b;
    }
    if (g.work_rows != g.in_rows || g.work_cols != g.in_cols) {
        ClassTimer t(ctx, st, B2S_K_PRE, 1);
        void *dst = (cur == s.pre_a) ? s.pre_b : s.pre_a;
        b2s_launch_block_reduce(cur, cur_dt, g.in_rows, g.in_cols, g.dsy, g.dsx, p.down_sample_method, dst, g.work_dtype,
                                g.work_rows, g.work_cols, nb, st);
        cur = dst;
        cur_dt = g.work_dtype;
    }

    B2sImg padded = img_of(pl, s.padded, 0);
    if (g.f64) {   // integer pixels, no log: float64 pad -> wavedec2 -> notch -> waverec2, rint / clip into `padded`
        ClassTimer t(ctx, st, B2S_K_OTHER, 2 + g.n_passes * (1 + g.levels * (6 + (p.bidirectional ? 2 : 1))));
        B2sF64Args f;
        memset(&f, 0, sizeof f);
        f.in = cur; f.in_dtype = cur_dt; f.rows = g.work_rows; f.cols = g.work_cols;
        f.pad_mode = p.pad_mode; f.base_pad = g.base_pad;
        f.pad_value = p.pad_mode == B2S_PAD_CONSTANT ? p.pad_constant : 0.0;
        f.PH = g.PH; f.PW = g.PW; f.levels = g.levels;
        for (int l = 0; l <= g.levels; ++l) { f.my[l] = g.my[l]; f.mx[l] = g.mx[l]; }
        f.F = pl->taps.F;
        f.dec_lo = pl->taps64[0]; f.dec_hi = pl->taps64[1]; f.rec_lo = pl->taps64[2]; f.rec_hi = pl->taps64[3];
        f.n_passes = g.n_passes; f.bidirectional = p.bidirectional;
        for (int pass = 0; pass < g.n_passes; ++pass)
            for (int l = 1; l <= g.levels; ++l)
                for (int ax = 0; ax < (p.bidirectional ? 2 : 1); ++ax) {
                    if (!pl->d_notch_mat[pass][l][ax])
                        return fail(ctx, B2S_ERR_INVALID, "float64 destripe: b2s_plan_set_notch_matrix was not called for pass %d, level %d, axis %d", pass, l, ax);
                    f.notch[pass][l][ax] = pl->d_notch_mat[pass][l][ax];
                }
        f.work = s.f64_work; f.plane_doubles = pl->f64_plane_doubles;
        f.out = padded;
        if (p.mask) {   // core.py:1079-1080 on the integer image
            ClassTimer tm(ctx, st, B2S_K_OTHER, 12);
            b2s_launch_mask_threshold_int(cur, cur_dt, (size_t)g.work_rows * g.work_cols, p.mask_threshold,
                                          p.mask_per_plane ? pl->d_mask_thr_pp + z0 : nullptr, s.mask_a, nb, st);
            const int rc = b2s_launch_img_mask(s.mask_a, s.mask_b, s.mask_c, g.work_rows, g.work_cols, p.mask_close, p.mask_open,
                                               s.mask_flag, s.mask_hflag, nb, st);
            if (rc) return fail(ctx, rc, "enable_masking: get_img_mask failed (rows of %d pixels)", g.work_cols);
            f.mask = s.mask_a;
        }
        b2s_launch_f64_destripe(f, nb, st);
    } else if (g.log_image) {
        {
            ClassTimer t(ctx, st, B2S_K_PROLOGUE, 1);
            B2sPrologueArgs a;
            a.in = cur;
            a.in_dtype = cur_dt;
            a.src_rows = g.work_rows;
            a.src_cols = g.work_cols;
            a.flat = g.fuse_flat ? pl->d_flat : nullptr;
            a.pad_mode = p.pad_mode;
            a.base_pad = g.base_pad;
            a.use_log1p = p.log1p;
            a.pad_value = p.pad_mode == B2S_PAD_CONSTANT ? (float)p.pad_constant : 0.f;
            a.pad_value_pp = (p.bleach_per_plane && p.pad_mode == B2S_PAD_CONSTANT && pl->d_padv_pp) ? pl->d_padv_pp + z0 : nullptr;
            a.out = padded;
            a.n_groups = pl->n_row_groups;
            a.row_src = pl->d_row_src;
            a.row_start = pl->d_row_start;
            a.row_targets = pl->d_row_targets;
            a.colmap = pl->d_colmap;
            a.lut = (cur_dt != B2S_F32 && !a.flat && p.log1p) ? pl->d_lut : nullptr;
            a.minmax = fuse_minmax ? s.mm : nullptr;
            b2s_launch_prologue(a, nb, st);
            if (p.mask) {   // core.py:1079-1080: img *= get_img_mask(img, clip_med) on the (log) image, ahead of numpy.pad
                ctx->launches += 12;
                b2s_launch_mask_threshold_f32(padded, g.base_pad, g.work_rows, g.work_cols, p.mask_threshold,
                                              p.mask_per_plane ? pl->d_mask_thr_pp + z0 : nullptr, s.mask_a, nb, st);
                const int rc = b2s_launch_img_mask(s.mask_a, s.mask_b, s.mask_c, g.work_rows, g.work_cols, p.mask_close, p.mask_open,
                                                   s.mask_flag, s.mask_hflag, nb, st);
                if (rc) return fail(ctx, rc, "enable_masking: get_img_mask failed (rows of %d pixels)", g.work_cols);
                b2s_launch_mask_apply(padded, s.mask_a, g.base_pad, g.work_rows, g.work_cols, p.pad_mode, nb, st);
            }
            if (g.n_passes > 0 && p.pad_mode >= B2S_PAD_LINEAR_RAMP) {   // pad areas computed from the image (numpy.pad stat modes)
                ctx->launches += p.pad_mode == B2S_PAD_EMPTY ? 0 : (p.pad_mode == B2S_PAD_LINEAR_RAMP ? 4 : 2);
                b2s_launch_pad_fill(p.pad_mode, padded, g.base_pad, g.work_rows, g.work_cols, s.pad_flags, nb, st);
            }
        }
        if (p.debug_stop_after == B2S_STAGE_PROLOGUE) return B2S_OK;
        for (int pass = 0; pass < g.n_passes; ++pass) {
            {
                for (int l = 1; l <= g.levels; ++l) {
                    ClassTimer t(ctx, st, B2S_K_DWT_FWD, 1, l);
                    const B2sImg in = l == 1 ? padded : img_of(pl, s.sub[l - 1][0], l - 1);
                    b2s_launch_dwt_fwd(pl->taps, in, img_of(pl, s.sub[l][0], l), img_of(pl, s.sub[l][1], l),
                                       img_of(pl, s.sub[l][2], l), img_of(pl, s.sub[l][3], l), nb, exact, ctx->sm_count, st,
                                       s.dwt_scratch, pl->dwt_scratch_stride);
                }
            }
            if (p.debug_stop_after == B2S_STAGE_FORWARD) return B2S_OK;
            {
                for (int l = 1; l <= g.levels; ++l) {
                    ClassTimer t(ctx, st, B2S_K_NOTCH, p.bidirectional ? 2 : 1, l);
                    auto notch = [&](int n, const float *gt, float *band, int along_cols) {
                        auto it = pl->xfft.find(n);
                        if (it != pl->xfft.end())
                            b2s_launch_notch_exact(it->second, gt, img_of(pl, band, l), along_cols, nb, ctx->sm_count, st);
                        else
                            b2s_launch_notch(pl->fft[n], gt, img_of(pl, band, l), along_cols, nb, ctx->sm_count, st);
                    };
                    notch(g.mx[l], pl->d_notch[pass][l][0], s.sub[l][1], 0);
                    if (p.bidirectional) notch(g.my[l], pl->d_notch[pass][l][1], s.sub[l][2], 1);
                }
            }
            if (p.debug_stop_after == B2S_STAGE_NOTCH) return B2S_OK;
            {
                for (int l = g.levels; l >= 1; --l) {
                    ClassTimer t(ctx, st, B2S_K_DWT_INV, 1, l);
                    // the reconstruction of level l-1 overwrites that level's approximation buffer (or the padded image)
                    B2sImg out = l == 1 ? padded : img_of(pl, s.sub[l - 1][0], l - 1);
                    b2s_launch_dwt_inv(pl->taps, img_of(pl, s.sub[l][0], l), img_of(pl, s.sub[l][1], l),
                                       img_of(pl, s.sub[l][2], l), img_of(pl, s.sub[l][3], l), out, nb, exact, ctx->sm_count, st,
                                       s.dwt_scratch, pl->dwt_scratch_stride);
                }
            }
        }
        if (p.debug_stop_after == B2S_STAGE_INVERSE) return B2S_OK;
        if (p.bleach) {   // core.py:1131-1139: after the crop, before expm1
            ClassTimer t(ctx, st, B2S_K_OTHER, 3);
            B2sBleachArgs b;
            b.img = padded;
            b.base_pad = g.base_pad; b.rows = g.work_rows; b.cols = g.work_cols;
            b.b0 = p.bleach_b0; b.b1 = p.bleach_b1; b.a1 = p.bleach_a1; b.zi = p.bleach_zi;
            b.clip_min = p.bleach_clip_min; b.clip_med = p.bleach_clip_med; b.clip_max = p.bleach_clip_max;
            b.clip_pp = p.bleach_per_plane ? pl->d_clip_pp + 3 * z0 : nullptr;
            b.scratch = s.bleach_scratch;
            b.scratch_plane_stride = (size_t)g.work_rows * (g.work_cols + 12);
            b.filt = s.bleach_filt;
            b.maxkey = s.bleach_max;
            if (p.bleach == 2) b2s_launch_bleach_max_method(b, nb, st);
            else b2s_launch_bleach(b, nb, st);
        }
    }
    {
        B2sEpilogueArgs e;
        memset(&e, 0, sizeof e);
        e.in = padded;
        e.raw = cur;
        e.raw_dtype = cur_dt;
        e.destripe = g.log_image;
        e.base_pad = g.base_pad;
        e.rows = g.work_rows;
        e.cols = g.work_cols;
        e.use_log1p = p.log1p;
        e.int_path = g.int_path;
        e.work_dtype = g.work_dtype;
        e.dark = p.process_img ? p.dark : 0.0;
        // float32 holds every intermediate exactly unless a fractional dark meets an integer image (float64 there)
        e.f32_exact = !(e.dark > 0.0 && g.int_path && e.dark != std::floor(e.dark)) && e.dark < 16777216.0;
        e.ls_sub = nullptr;
        e.final_mode = g.final_mode;
        e.shift = p.bit_shift_to_right;
        e.out_dtype = g.out_dtype;
        e.flip = p.process_img ? p.flip_upside_down : 0;
        e.rot = p.process_img ? p.rotate / 90 : 0;
        e.uniform_mm = p.process_img ? s.mm : nullptr;
        e.out = d_out;
        e.out_rows = g.out_rows;
        e.out_cols = g.out_cols;
        e.epi_thr = pl->d_epi_thr;
        e.epi_kmax = pl->epi_kmax;
        // the image as the reference holds it after the dark subtraction (same dtype), unrotated
        B2sEpilogueArgs m = e;
        const bool mid_int = g.work_dtype != B2S_F32;
        m.final_mode = mid_int ? 0 : 3;
        m.out_dtype = mid_int ? g.work_dtype : B2S_F32;
        m.flip = 0; m.rot = 0; m.uniform_mm = nullptr;
        m.out = s.mid; m.out_rows = g.work_rows; m.out_cols = g.work_cols;
        if (g.resize) {
            {
                ClassTimer t(ctx, st, B2S_K_EPILOGUE, 1);
                b2s_launch_epilogue(m, nb, st);
            }
            const void *src = s.mid;
            if (pl->ls) {   // lightsheet clean into a second work-size image of d_type, then resize that
                ClassTimer t2(ctx, st, B2S_K_LIGHTSHEET, s.ls_cells ? 4 : 3);
                B2sEpilogueArgs le = e;
                le.final_mode = 0; le.out_dtype = g.mid_dtype; le.flip = 0; le.rot = 0;
                le.out = s.mid2; le.out_rows = g.work_rows; le.out_cols = g.work_cols;
                b2s_launch_lightsheet(pl->ls, s.mid, s.ls_grid, s.bg_grid, s.ls_cells, le, nb, st);
                src = s.mid2;
            }
            ClassTimer t3(ctx, st, B2S_K_EPILOGUE, 2);
            CU(ctx, cudaMemsetAsync(s.mm2, 0xff, sizeof(unsigned) * 2 * nb, st));
            b2s_launch_minmax(src, g.mid_dtype, (size_t)g.work_rows * g.work_cols, nb, s.mm2, st);
            int rdt = g.mid_dtype;
            if (g.aa) {   // scipy.ndimage.gaussian_filter: axis 0, then axis 1, each pass stored in the output dtype
                const int f64 = g.mid_dtype != B2S_F32;
                void *bufs[2] = {s.aa_a, s.aa_b};
                const int radius[2] = {p.aa_radius_y, p.aa_radius_x};
                int k = 0;
                for (int axis = 0; axis < 2; ++axis) {
                    if (radius[axis] <= 0) continue;
                    if (!pl->d_aa_w[axis]) return fail(ctx, B2S_ERR_INVALID, "aa_radius is set but b2s_plan_set_aa_weights was not called for axis %d", axis);
                    ctx->launches += 1;
                    b2s_launch_gauss_aa(src, rdt, bufs[k], f64, g.work_rows, g.work_cols, axis, pl->d_aa_w[axis], radius[axis], nb, st);
                    src = bufs[k++];
                    rdt = f64 ? B2S_F64_INTERNAL : B2S_F32;
                }
            }
            B2sResizeArgs r;
            r.src = src; r.dtype = rdt; r.mm_dtype = g.mid_dtype; r.rows = g.work_rows; r.cols = g.work_cols;
            r.new_rows = g.new_rows; r.new_cols = g.new_cols;
            r.iy0 = pl->d_rz_idx; r.iy1 = r.iy0 + g.new_rows; r.ix0 = r.iy1 + g.new_rows; r.ix1 = r.ix0 + g.new_cols;
            r.wy0 = pl->d_rz_w; r.wy1 = r.wy0 + g.new_rows; r.wx0 = r.wy1 + g.new_rows; r.wx1 = r.wx0 + g.new_cols;
            r.mm = s.mm2;
            b2s_launch_resize_final(r, e, nb, st);
        } else if (!pl->ls) {
            ClassTimer t(ctx, st, B2S_K_EPILOGUE, 1);
            b2s_launch_epilogue(e, nb, st);
        } else {
            // stage 1: the post-dark image
            {
                ClassTimer t(ctx, st, B2S_K_EPILOGUE, 1);
                b2s_launch_epilogue(m, nb, st);
            }
            // stage 2: percentile grids, zoom, subtraction, final conversion
            ClassTimer t2(ctx, st, B2S_K_LIGHTSHEET, s.ls_cells ? 4 : 3);
            b2s_launch_lightsheet(pl->ls, s.mid, s.ls_grid, s.bg_grid, s.ls_cells, e, nb, st);
        }
    }
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

// Page-locked host memory on the NUMA node the GPU hangs off: the pages are faulted in by the calling thread, so the
// thread is confined to the GPU's local CPUs (sysfs local_cpulist of its PCI function) for the duration of the
// allocation and its affinity restored afterwards.  With several GPUs per box (one process each) this keeps every
// H2D / D2H stream on its own socket's memory controllers.  B2S_NUMA=0 disables it; any failure falls back to a plain
// cudaMallocHost.
cudaError_t numa_local_malloc_host(b2s_context *ctx, void **ptr, size_t bytes)
{
    static const int enabled = getenv("B2S_NUMA") ? atoi(getenv("B2S_NUMA")) : 1;
    cpu_set_t saved, local;
    bool bound = false;
    if (enabled && sched_getaffinity(0, sizeof saved, &saved) == 0) {
        char bus[32] = {0}, path[128];
        if (cudaDeviceGetPCIBusId(bus, sizeof bus, ctx->device) == cudaSuccess) {
            for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
            snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
            if (FILE *f = fopen(path, "r")) {
                char line[4096] = {0};
                if (fgets(line, sizeof line, f)) {
                    CPU_ZERO(&local);
                    int n_local = 0;
                    char *save = nullptr;   // strtok_r: feeder threads of several GPUs allocate concurrently
                    for (char *tok = strtok_r(line, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
                        int a = 0, b = 0;
                        const int k = sscanf(tok, "%d-%d", &a, &b);
                        if (k == 1) b = a;
                        if (k >= 1)
                            for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
                                if (CPU_ISSET(c, &saved)) { CPU_SET(c, &local); ++n_local; }
                    }
                    if (n_local > 0 && sched_setaffinity(0, sizeof local, &local) == 0) bound = true;
                }
                fclose(f);
            }
        } else {
            cudaGetLastError();
        }
    }
    const cudaError_t e = cudaMallocHost(ptr, bytes);
    if (bound) sched_setaffinity(0, sizeof saved, &saved);
    return e;
}

bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// worker threads for pageable staging: B2S_HOST_THREADS, else min(8, cores / (2 x processes on this box))
HostPool *host_pool(b2s_context *ctx)
{
    static std::mutex create_mutex;   // plans of one context may run from several host threads (each under its own plan lock)
    std::lock_guard<std::mutex> lk(create_mutex);
    if (!ctx->pool) {
        int n = getenv("B2S_HOST_THREADS") ? atoi(getenv("B2S_HOST_THREADS")) : 0;
        if (n <= 0) {
            const char *w = getenv("LOCAL_WORLD_SIZE") ? getenv("LOCAL_WORLD_SIZE") : getenv("WORLD_SIZE");
            const int world = std::max(1, w ? atoi(w) : 1);
            const int hw = (int)std::thread::hardware_concurrency();
            n = std::max(1, std::min(8, hw / (2 * world)));
        }
        ctx->pool = new HostPool(n);
    }
    return ctx->pool;
}

// Staging copy with non-temporal stores: the destination (a pinned staging buffer, or the caller's result array) is not read
// again by this core, so write-allocate traffic — one extra DRAM read per written line with ordinary stores below glibc's
// non-temporal threshold — is pure loss on a path bounded by host memory bandwidth.  B2S_HOST_NT=0 restores memcpy.
void stream_copy(void *dst, const void *src, size_t n)
{
    static const bool nt = !(getenv("B2S_HOST_NT") && atoi(getenv("B2S_HOST_NT")) == 0);
    if (!nt || n < 4096) { memcpy(dst, src, n); return; }
    char *d = (char *)dst;
    const char *s = (const char *)src;
    const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(s)), b = _mm_loadu_si128((const __m128i *)(s + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
        _mm_stream_si128((__m128i *)(d), a); _mm_stream_si128((__m128i *)(d + 16), b);
        _mm_stream_si128((__m128i *)(d + 32), c); _mm_stream_si128((__m128i *)(d + 48), e);
        s += 64; d += 64;
    }
    _mm_sfence();
    if (n & 63) memcpy(d, s, n & 63);
}

// memcpy split over the pool; `grp` counts the chunks still running
void copy_async(HostPool *pool, TaskGroup *grp, void *dst, const void *src, size_t bytes)
{
    const size_t min_chunk = 1 << 20;
    int parts = (int)std::min<size_t>((size_t)pool->size(), (bytes + min_chunk - 1) / min_chunk);
    if (parts < 1) parts = 1;
    const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
    grp->add(parts);
    for (int i = 0; i < parts; ++i) {
        const size_t o = std::min(bytes, per * i), e = std::min(bytes, per * (i + 1));
        pool->submit([=] { if (e > o) stream_copy((char *)dst + o, (const char *)src + o, e - o); grp->done(); });
    }
}

}  // namespace

extern "C" {

int b2s_version(void) { return B2S_VERSION; }

void b2s_params_default(b2s_params *p)
{
    memset(p, 0, sizeof *p);
    p->struct_size = (int32_t)sizeof *p;
    p->in_dtype = p->out_dtype = B2S_U16;
    p->pad_mode = B2S_PAD_WRAP;
    p->log1p = 1;
    p->artifact_length = 150;
    p->background_window_size = 200;
    p->percentile = 0.25;
    p->lightsheet_vs_background = 2.0;
    p->bit_shift_to_right = 8;
    p->max_batch = 8;
    p->exact = 1;
}

int b2s_create(int device, b2s_context **out)
{
    if (!out) return B2S_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    b2s_context *ctx = new b2s_context();
    *out = ctx;  // returned even on failure so that b2s_last_error works; caller destroys it
    if (e != cudaSuccess || n == 0)
        return fail(ctx, B2S_ERR_CUDA, "no CUDA device available (%s) — this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(ctx, B2S_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
    ctx->device = device;
    CU(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10)
        return fail(ctx, B2S_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    {   // temporaries of the stream-ordered entries (histogram, mask, deflate) stay in the pool between calls (up to 1 GiB)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return B2S_OK;
}

void b2s_destroy(b2s_context *ctx)
{
    if (!ctx) return;
    for (auto &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    delete ctx->pool;
    delete ctx;
}

const char *b2s_last_error(const b2s_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int b2s_device_sm_count(const b2s_context *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t b2s_launch_count(const b2s_context *ctx) { return ctx ? ctx->launches : 0; }

int b2s_plan_geometry(const b2s_params *params, b2s_plan_info *info, char *err, size_t err_len)
{
    if (!params || !info) return B2S_ERR_INVALID;
    b2s_context tmp;
    Geometry g;
    int rc = compute_geometry(&tmp, *params, g);
    if (rc) {
        if (err && err_len) snprintf(err, err_len, "%s", tmp.err.c_str());
        return rc;
    }
    fill_info(*params, g, info);
    return B2S_OK;
}

int b2s_resize_table(int n_in, int n_out, int32_t *idx0, int32_t *idx1, double *w0, double *w1)
{
    if (n_in <= 0 || n_out <= 0 || !idx0 || !idx1 || !w0 || !w1) return B2S_ERR_INVALID;
    b2s_resize_axis_table(n_in, n_out, idx0, idx1, w0, w1);
    return B2S_OK;
}

int b2s_plan_create(b2s_context *ctx, const b2s_params *params, b2s_plan **out)
{
    if (!ctx || !params || !out) return B2S_ERR_INVALID;
    *out = nullptr;
    Geometry g;
    int rc = compute_geometry(ctx, *params, g);
    if (rc) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    b2s_plan *pl = new b2s_plan();
    pl->ctx = ctx;
    pl->p = *params;
    pl->g = g;
    pl->B = params->max_batch > 0 ? params->max_batch : 8;
    for (int l = 0; l <= g.levels; ++l) {
        pl->pitch[l] = round4(g.mx[l]);
        pl->plane_stride[l] = (size_t)pl->pitch[l] * g.my[l];
    }
    if (g.n_passes == 0) { pl->pitch[0] = round4(g.PW); pl->plane_stride[0] = (size_t)pl->pitch[0] * g.PH; }
    if (g.log_image) {
        const int need = g.n_passes > 0 ? b2s_dwt_max_smem(params->n_taps) : 0;
        if (need > 227 * 1024) { delete pl; return fail(ctx, B2S_ERR_UNSUPPORTED, "filter too long for the shared-memory tiles"); }
        if (g.n_passes == 0) pl->p.n_taps = 0;     // no filter bank: build_tables only lays out the prologue maps
        rc = build_tables(pl);
    }
    if (rc == B2S_OK) rc = build_resize_tables(pl);
    if (rc == B2S_OK && params->process_img && params->lightsheet) {
        pl->ls = b2s_lightsheet_create(g.work_rows, g.work_cols, g.work_dtype, params->artifact_length,
                                       params->background_window_size, params->percentile,
                                       params->lightsheet_vs_background, 1);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B2S_ERR_CUDA, "lightsheet table setup failed");
    }
    if (rc == B2S_OK) rc = fit_workspace(pl);
    for (int si = 0; rc == B2S_OK && si < pl->n_slots; ++si) rc = alloc_slot(pl, si);
    if (rc == B2S_OK && cudaEventCreateWithFlags(&pl->fork, cudaEventDisableTiming) != cudaSuccess) rc = fail(ctx, B2S_ERR_CUDA, "cudaEventCreate failed");
    pl->p.dec_lo = nullptr;  // the caller's table is not retained
    if (rc) { b2s_plan_destroy(pl); return rc; }
    *out = pl;
    return B2S_OK;
}

void b2s_plan_destroy(b2s_plan *pl)
{
    if (!pl) return;
    cudaSetDevice(pl->ctx->device);
    cudaDeviceSynchronize();
    for (void *p : pl->allocs) cudaFree(p);
    for (auto &kv : pl->xfft) b2s_xfft_destroy(kv.second);
    b2s_lightsheet_destroy(pl->ls);
    if (pl->fork) cudaEventDestroy(pl->fork);
    for (auto &s : pl->slot) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.mask_hflag) cudaFreeHost(s.mask_hflag);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
    }
    delete pl;
}

int b2s_plan_query(const b2s_plan *pl, b2s_plan_info *info)
{
    if (!pl || !info) return B2S_ERR_INVALID;
    b2s_params p = pl->p;
    fill_info(p, pl->g, info);
    info->workspace_bytes = pl->workspace_bytes;
    return B2S_OK;
}

int b2s_plan_set_flat(b2s_plan *pl, const float *flat, int is_device)
{
    if (!pl || !flat) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(float) * (size_t)pl->g.in_rows * pl->g.in_cols;
    if (!pl->d_flat) {
        int rc = dev_alloc(pl, (void **)&pl->d_flat, bytes);
        if (rc) return rc;
    }
    CU(ctx, cudaMemcpy(pl->d_flat, flat, bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    return B2S_OK;
}

int b2s_plan_set_aa_weights(b2s_plan *pl, int axis, const double *w, int n)
{
    if (!pl || !w) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    if (axis < 0 || axis > 1) return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_aa_weights: axis must be 0 (rows) or 1 (columns)");
    const int r = axis == 0 ? pl->p.aa_radius_y : pl->p.aa_radius_x;
    if (!pl->g.aa || r <= 0 || n != 2 * r + 1)
        return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_aa_weights: the plan expects %d weights along axis %d, got %d", r > 0 ? 2 * r + 1 : 0, axis, n);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    if (!pl->d_aa_w[axis]) {
        int rc = dev_alloc(pl, (void **)&pl->d_aa_w[axis], sizeof(double) * n);
        if (rc) return rc;
    }
    CU(ctx, cudaMemcpy(pl->d_aa_w[axis], w, sizeof(double) * n, cudaMemcpyHostToDevice));
    return B2S_OK;
}

int b2s_plan_wants_notch_matrix(const b2s_plan *pl) { return pl && pl->g.f64 ? 1 : 0; }

int b2s_plan_set_notch_matrix(b2s_plan *pl, int pass, int level, int axis, const double *R, int n)
{
    if (!pl || !R) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    const Geometry &gm = pl->g;
    if (!gm.f64) return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch_matrix: the plan does not run in float64");
    if (pass < 0 || pass >= gm.n_passes || level < 1 || level > gm.levels || axis < 0 || axis > (pl->p.bidirectional ? 1 : 0))
        return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch_matrix: no such matrix (pass %d, level %d, axis %d)", pass, level, axis);
    const int len = axis == 0 ? gm.mx[level] : gm.my[level];
    if (n != len) return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch_matrix: %d x %d given, sub-band side is %d", n, n, len);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    if (!pl->d_notch_mat[pass][level][axis]) {
        int rc = dev_alloc(pl, (void **)&pl->d_notch_mat[pass][level][axis], sizeof(double) * (size_t)n * n);
        if (rc) return rc;
    }
    CU(ctx, cudaMemcpy(pl->d_notch_mat[pass][level][axis], R, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    return B2S_OK;
}

int b2s_plan_set_bleach_levels(b2s_plan *pl, const double *clip, const float *pad_value, int64_t n_planes)
{
    if (!pl || !clip || n_planes <= 0) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    if (!pl->p.bleach || !pl->p.bleach_per_plane)
        return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_bleach_levels: the plan was not created with bleach_per_plane");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    if (n_planes > pl->cap_levels) {   // grow (the old blocks stay in the plan's allocation list until it is destroyed)
        int rc = dev_alloc(pl, (void **)&pl->d_clip_pp, sizeof(double) * 3 * (size_t)n_planes);
        if (rc) return rc;
        if ((rc = dev_alloc(pl, (void **)&pl->d_padv_pp, sizeof(float) * (size_t)n_planes))) return rc;
        pl->cap_levels = n_planes;
    }
    CU(ctx, cudaMemcpy(pl->d_clip_pp, clip, sizeof(double) * 3 * (size_t)n_planes, cudaMemcpyHostToDevice));
    if (pad_value) CU(ctx, cudaMemcpy(pl->d_padv_pp, pad_value, sizeof(float) * (size_t)n_planes, cudaMemcpyHostToDevice));
    else CU(ctx, cudaMemset(pl->d_padv_pp, 0, sizeof(float) * (size_t)n_planes));
    pl->n_levels = n_planes;
    return B2S_OK;
}

int b2s_plan_set_mask_thresholds(b2s_plan *pl, const double *thr, int64_t n_planes)
{
    if (!pl || !thr || n_planes <= 0) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    if (!pl->p.mask || !pl->p.mask_per_plane)
        return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_mask_thresholds: the plan was not created with mask_per_plane");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    if (n_planes > pl->cap_mask_thr) {
        int rc = dev_alloc(pl, (void **)&pl->d_mask_thr_pp, sizeof(double) * (size_t)n_planes);
        if (rc) return rc;
        pl->cap_mask_thr = n_planes;
    }
    CU(ctx, cudaMemcpy(pl->d_mask_thr_pp, thr, sizeof(double) * (size_t)n_planes, cudaMemcpyHostToDevice));
    pl->n_mask_thr = n_planes;
    return B2S_OK;
}

int b2s_plan_set_notch(b2s_plan *pl, int pass, int level, int axis, const float *g, int n)
{
    if (!pl || !g) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    const Geometry &gm = pl->g;
    if (pass < 0 || pass >= gm.n_passes || level < 1 || level > gm.levels || axis < 0 || axis > (pl->p.bidirectional ? 1 : 0))
        return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch: no such table (pass %d, level %d, axis %d)", pass, level, axis);
    const int len = axis == 0 ? gm.mx[level] : gm.my[level];
    if (n != len) return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch: table has %d entries, sub-band side is %d", n, len);
    if (gm.f64) return fail(ctx, B2S_ERR_INVALID, "b2s_plan_set_notch: a float64 plan takes b2s_plan_set_notch_matrix");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(pl->d_notch[pass][level][axis], g, sizeof(float) * n, cudaMemcpyHostToDevice));
    return B2S_OK;
}

int b2s_run(b2s_plan *pl, const void *in, void *out, int64_t n_planes, int in_is_device, int out_is_device, void *stream)
{
    if (!pl || !in || !out || n_planes < 0) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    const Geometry &g = pl->g;
    const size_t in_plane = (size_t)g.in_rows * g.in_cols * dtype_size(pl->p.in_dtype);
    const size_t out_plane = (size_t)g.out_rows * g.out_cols * dtype_size(g.out_dtype);
    const int B = pl->B;
    if (pl->p.bleach && pl->p.bleach_per_plane && pl->n_levels < n_planes)
        return fail(ctx, B2S_ERR_INVALID, "bleach_per_plane: b2s_plan_set_bleach_levels supplied %lld planes, the run has %lld",
                    (long long)pl->n_levels, (long long)n_planes);
    if (pl->p.mask && pl->p.mask_per_plane && g.log_image && pl->n_mask_thr < n_planes)
        return fail(ctx, B2S_ERR_INVALID, "mask_per_plane: b2s_plan_set_mask_thresholds supplied %lld planes, the run has %lld",
                    (long long)pl->n_mask_thr, (long long)n_planes);

    if (in_is_device && out_is_device) {
        cudaStream_t st = (cudaStream_t)stream;
        // more than one batch: consecutive batches alternate between slots on the slots' own streams, forked from and
        // joined back into the caller's stream, so the small launches of the deep levels of one batch (grids below one
        // wave) overlap the large launches of the next.  Per-kernel timing keeps everything on the caller's stream.
        static const int dev_slots_env = getenv("B2S_DEV_SLOTS") ? atoi(getenv("B2S_DEV_SLOTS")) : 3;
        const int n_batches = (int)((n_planes + B - 1) / B);
        const int ns = ctx->timing ? 1 : std::max(1, std::min(std::min(dev_slots_env, pl->n_slots), n_batches));
        if (ns > 1) {
            CU(ctx, cudaEventRecord(pl->fork, st));
            for (int k = 0; k < ns; ++k) CU(ctx, cudaStreamWaitEvent(pl->slot[k].stream, pl->fork, 0));
        }
        int si = 0;
        for (int64_t z = 0; z < n_planes; z += B) {
            const int nb = (int)std::min<int64_t>(B, n_planes - z);
            b2s_plan::Slot &s = pl->slot[si];
            int rc = enqueue_batch(pl, s, (const char *)in + z * in_plane, (char *)out + z * out_plane, nb, ns > 1 ? s.stream : st, z);
            if (rc) return rc;
            si = (si + 1) % ns;
        }
        if (ns > 1) {
            for (int k = 0; k < ns; ++k) {
                CU(ctx, cudaEventRecord(pl->slot[k].done, pl->slot[k].stream));
                CU(ctx, cudaStreamWaitEvent(st, pl->slot[k].done, 0));
            }
        }
        return B2S_OK;
    }

    // host path: kSlots slots, each with its own stream: H2D -> kernels -> D2H; the slots overlap each other
    const bool in_pinned = in_is_device || is_pinned_host(in);
    const bool out_pinned = out_is_device || is_pinned_host(out);
    for (int k = 0; k < pl->n_slots; ++k) {
        b2s_plan::Slot &s = pl->slot[k];
        if (!in_pinned && !s.h_in) CU(ctx, numa_local_malloc_host(ctx, &s.h_in, in_plane * B));
        if (!out_pinned && !s.h_out) CU(ctx, numa_local_malloc_host(ctx, &s.h_out, out_plane * B));
    }
    struct Pending { int64_t z; int nb; bool active; } pend[b2s_plan::kSlots] = {};
    // pageable buffers: the staging copies run on the context's worker threads.  Per slot, `in_grp` counts the chunks of
    // the running copy caller -> h_in, `out_grp` those of h_out -> caller; both directions proceed at the same time.
    HostPool *pool = (!in_pinned || !out_pinned) ? host_pool(ctx) : nullptr;
    TaskGroup in_grp[b2s_plan::kSlots], out_grp[b2s_plan::kSlots];
    struct WaitAll {   // no early return may leave worker threads writing into the caller's buffers
        TaskGroup *a, *b; int n;
        ~WaitAll() { for (int i = 0; i < n; ++i) { a[i].wait(); b[i].wait(); } }
    } wait_all{in_grp, out_grp, b2s_plan::kSlots};
    auto drain = [&](int si) -> int {
        if (!pend[si].active) return B2S_OK;
        b2s_plan::Slot &s = pl->slot[si];
        CU(ctx, cudaEventSynchronize(s.done));
        if (!out_is_device && !out_pinned)
            copy_async(pool, &out_grp[si], (char *)out + pend[si].z * out_plane, s.h_out, out_plane * pend[si].nb);
        pend[si].active = false;
        return B2S_OK;
    };
    // batch sizes ramp up from 4 planes and down again at the end of the call: the first kernels start after a short
    // copy and the last device-to-host copy is short, so the pipeline fill / drain costs little
    // the host path runs shorter batches than the workspace allows: finer pipelining of H2D / kernels / D2H across the
    // slots hides more of the PCIe time (measured at 2048^2: 8 planes per batch, 21.5 Gpx/s end to end vs 17.7 with 32)
    static const int host_cap = getenv("B2S_HOST_BATCH") ? atoi(getenv("B2S_HOST_BATCH")) : 8;
    const int Bh = std::max(1, std::min(B, host_cap));
    static const int host_slots_env = getenv("B2S_HOST_SLOTS") ? atoi(getenv("B2S_HOST_SLOTS")) : b2s_plan::kSlots;
    const int nsh = std::max(1, std::min(pl->n_slots, host_slots_env));
    int si = 0;
    int64_t z = 0;
    static const int ramp_env = std::max(1, getenv("B2S_HOST_RAMP") ? atoi(getenv("B2S_HOST_RAMP")) : 4);   // first / last batch size
    const int ramp = std::min(ramp_env, Bh);   // never above the batch the slot buffers were sized for
    int nb_next = ramp;
    while (z < n_planes) {
        const int64_t left = n_planes - z;
        int nb = (int)std::min<int64_t>(nb_next, left);
        if (left > ramp && nb > left / 2) nb = (int)std::max<int64_t>(ramp, left / 2);   // ramp down: halve what is left
        nb = std::max(1, std::min(nb, Bh));   // every slot buffer holds exactly B planes (B = 3 with 5, 8, 11 ... planes left)
        nb_next = std::min(Bh, nb_next * 2);
        b2s_plan::Slot &s = pl->slot[si];
        int rc = drain(si);   // the slot's previous batch has left the device; its copy-out (if any) is now running
        if (rc) return rc;
        const void *d_in;
        if (in_is_device) {
            d_in = (const char *)in + z * in_plane;
        } else {
            const void *src = (const char *)in + z * in_plane;
            if (!in_pinned) {
                copy_async(pool, &in_grp[si], s.h_in, src, in_plane * nb);
                in_grp[si].wait();
                src = s.h_in;
            }
            CU(ctx, cudaMemcpyAsync(s.d_in, src, in_plane * nb, cudaMemcpyHostToDevice, s.stream));
            d_in = s.d_in;
        }
        void *d_out = out_is_device ? (void *)((char *)out + z * out_plane) : s.d_out;
        rc = enqueue_batch(pl, s, d_in, d_out, nb, s.stream, z);
        if (rc) return rc;
        if (!out_is_device) {
            void *dst = out_pinned ? (void *)((char *)out + z * out_plane) : s.h_out;
            if (!out_pinned) out_grp[si].wait();   // h_out is free once the previous batch has been copied out of it
            CU(ctx, cudaMemcpyAsync(dst, s.d_out, out_plane * nb, cudaMemcpyDeviceToHost, s.stream));
        }
        CU(ctx, cudaEventRecord(s.done, s.stream));
        pend[si] = {z, nb, true};
        z += nb;
        si = (si + 1) % nsh;
    }
    for (int k = 0; k < pl->n_slots; ++k) {
        int rc = drain(k);
        if (rc) return rc;
    }
    return B2S_OK;   // ~WaitAll: the last copy-outs have landed in `out`
}

int b2s_host_alloc(b2s_context *ctx, size_t bytes, void **ptr)
{
    if (!ctx || !ptr) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = numa_local_malloc_host(ctx, ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(ctx, B2S_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return B2S_OK;
}

int b2s_host_free(b2s_context *ctx, void *ptr)
{
    if (!ctx) return B2S_ERR_INVALID;
    CU(ctx, cudaFreeHost(ptr));
    return B2S_OK;
}

int b2s_timing_enable(b2s_context *ctx, int on)
{
    if (!ctx) return B2S_ERR_INVALID;
    ctx->timing = on != 0;
    return B2S_OK;
}

int b2s_timing_read(b2s_context *ctx, double *ms, int64_t *launches, int reset)
{
    if (!ctx) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    for (auto &s : ctx->spans) {
        float t = 0.f;
        cudaEventElapsedTime(&t, s.a, s.b);
        const int lv = s.level < B2S_TIMING_LEVELS ? s.level : 0;
        ctx->ms[s.cls][lv] += t;
        ctx->n_launch[s.cls][lv] += s.n;
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    ctx->spans.clear();
    for (int k = 0; k < B2S_N_KERNEL_CLASSES; ++k)
        for (int l = 0; l < B2S_TIMING_LEVELS; ++l) {
            if (ms) ms[k * B2S_TIMING_LEVELS + l] = ctx->ms[k][l];
            if (launches) launches[k * B2S_TIMING_LEVELS + l] = ctx->n_launch[k][l];
            if (reset) { ctx->ms[k][l] = 0; ctx->n_launch[k][l] = 0; }
        }
    return B2S_OK;
}

int b2s_debug_read(b2s_plan *pl, int what, int level, int plane, float *out, int32_t *rows, int32_t *cols)
{
    if (!pl || !out) return B2S_ERR_INVALID;
    b2s_context *ctx = pl->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    const Geometry &g = pl->g;
    if (g.n_passes == 0) return fail(ctx, B2S_ERR_INVALID, "plan has no destripe workspace");
    if (plane < 0 || plane >= pl->B) return fail(ctx, B2S_ERR_INVALID, "plane out of range");
    int l = 0;
    const float *base;
    if (what == 0) base = pl->slot[0].padded;
    else {
        if (what < 1 || what > 4 || level < 1 || level > g.levels) return fail(ctx, B2S_ERR_INVALID, "bad sub-band / level");
        l = level;
        base = pl->slot[0].sub[l][what - 1];
    }
    const int r = g.my[l], c = g.mx[l];
    CU(ctx, cudaMemcpy2D(out, sizeof(float) * c, base + (size_t)plane * pl->plane_stride[l], sizeof(float) * pl->pitch[l],
                         sizeof(float) * c, r, cudaMemcpyDeviceToHost));
    if (rows) *rows = r;
    if (cols) *cols = c;
    return B2S_OK;
}

int b2s_isotropic_xy(b2s_context *ctx, const void *d_in, int in_dtype, int rows, int cols, int n_steps, const int32_t *steps,
                     int target_rows, int target_cols, int pre_rows, int pre_cols, const double *wy, int ry,
                     const double *wx, int rx, float *d_out, int n_planes, void *stream)
{
    if (!ctx || !d_in || !d_out || rows <= 0 || cols <= 0 || target_rows <= 0 || target_cols <= 0 || n_planes <= 0 ||
        n_steps < 0 || (n_steps && !steps) || in_dtype < B2S_U8 || in_dtype > B2S_F32 || ry < 0 || rx < 0 ||
        (ry && !wy) || (rx && !wx))
        return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t full = (size_t)rows * cols * n_planes;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (cudaMallocAsync(&p, bytes ? bytes : 16, st) != cudaSuccess) return nullptr;
        tmp.push_back(p);
        return p;
    };
    auto release = [&]() { for (void *p : tmp) cudaFreeAsync(p, st); tmp.clear(); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};   // early CU(...) returns free the temporaries too
    unsigned *mm0 = (unsigned *)dalloc(sizeof(unsigned) * 2 * n_planes);      // uniform check on the source plane
    unsigned *mm1 = (unsigned *)dalloc(sizeof(unsigned) * 2 * n_planes);      // clip range of the image resize() sees
    float *buf[2] = {(float *)dalloc(sizeof(float) * full), (float *)dalloc(sizeof(float) * full)};
    if (!mm0 || !mm1 || !buf[0] || !buf[1]) { release(); return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed: %s", cudaGetErrorString(cudaGetLastError())); }
    CU(ctx, cudaMemsetAsync(mm0, 0xff, sizeof(unsigned) * 2 * n_planes, st));
    CU(ctx, cudaMemsetAsync(mm1, 0xff, sizeof(unsigned) * 2 * n_planes, st));
    b2s_launch_minmax(d_in, in_dtype, (size_t)rows * cols, n_planes, mm0, st);
    // img.astype(float32) (an identity block reduce), then the alternating 2 x 1 / 1 x 2 reductions
    const void *cur = d_in;
    int cur_dt = in_dtype, r = rows, c = cols, k = 0, launches = 1;
    auto reduce = [&](int by, int bx, int method) {
        const int nr = (r + by - 1) / by, nc = (c + bx - 1) / bx;
        b2s_launch_block_reduce(cur, cur_dt, r, c, by, bx, method, buf[k], B2S_F32, nr, nc, n_planes, st);
        cur = buf[k]; cur_dt = B2S_F32; r = nr; c = nc; k ^= 1; ++launches;
    };
    // (the first reduction reads the integer plane itself: max / mean of two integers in float32 equals the same
    // operation on their float32 copies; only a plane that is not reduced at all needs the explicit conversion)
    for (int i = 0; i < n_steps; ++i) {
        const int ym = steps[2 * i], xm = steps[2 * i + 1];
        if (ym >= 0 && (r + 1) / 2 >= target_rows) reduce(2, 1, ym);
        if (xm >= 0 && (c + 1) / 2 >= target_cols) reduce(1, 2, xm);
    }
    if (cur_dt != B2S_F32) reduce(1, 1, B2S_DS_MAX);
    if (r != pre_rows || c != pre_cols) {
        release();
        return fail(ctx, B2S_ERR_INVALID, "isotropic down-sampling: the reductions end at %d x %d, the caller expected %d x %d", r, c, pre_rows, pre_cols);
    }
    b2s_launch_minmax(cur, B2S_F32, (size_t)r * c, n_planes, mm1, st);
    ++launches;
    // resize(anti_aliasing=True): Gaussian per axis (float32 image: each pass stored as float32), order-1 zoom, clip
    const int radius[2] = {ry, rx};
    const double *w[2] = {wy, wx};
    for (int axis = 0; axis < 2; ++axis) {
        if (radius[axis] <= 0) continue;
        const int n = 2 * radius[axis] + 1;
        double *dw = (double *)dalloc(sizeof(double) * n);
        if (!dw) { release(); return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed"); }
        CU(ctx, cudaMemcpyAsync(dw, w[axis], sizeof(double) * n, cudaMemcpyHostToDevice, st));
        b2s_launch_gauss_aa(cur, B2S_F32, buf[k], 0, r, c, axis, dw, radius[axis], n_planes, st);
        cur = buf[k]; k ^= 1; ++launches;
    }
    std::vector<int> idx(2 * (size_t)(target_rows + target_cols));
    std::vector<double> wt(2 * (size_t)(target_rows + target_cols));
    b2s_resize_axis_table(r, target_rows, idx.data(), idx.data() + target_rows, wt.data(), wt.data() + target_rows);
    b2s_resize_axis_table(c, target_cols, idx.data() + 2 * target_rows, idx.data() + 2 * target_rows + target_cols,
                          wt.data() + 2 * target_rows, wt.data() + 2 * target_rows + target_cols);
    int *d_idx = (int *)dalloc(sizeof(int) * idx.size());
    double *d_w = (double *)dalloc(sizeof(double) * wt.size());
    if (!d_idx || !d_w) { release(); return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed"); }
    CU(ctx, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_w, wt.data(), sizeof(double) * wt.size(), cudaMemcpyHostToDevice, st));
    B2sResizeArgs ra;
    ra.src = cur; ra.dtype = B2S_F32; ra.mm_dtype = B2S_F32; ra.rows = r; ra.cols = c;
    ra.new_rows = target_rows; ra.new_cols = target_cols;
    ra.iy0 = d_idx; ra.iy1 = ra.iy0 + target_rows; ra.ix0 = ra.iy1 + target_rows; ra.ix1 = ra.ix0 + target_cols;
    ra.wy0 = d_w; ra.wy1 = ra.wy0 + target_rows; ra.wx0 = ra.wy1 + target_rows; ra.wx1 = ra.wx0 + target_cols;
    ra.mm = mm1;
    B2sEpilogueArgs e;
    memset(&e, 0, sizeof e);
    e.final_mode = 3; e.out_dtype = B2S_F32; e.out = d_out; e.out_rows = target_rows; e.out_cols = target_cols;
    e.uniform_mm = mm0;                       // is_uniform_2d(img): zeros (parallel_image_processor.py:373-374)
    b2s_launch_resize_final(ra, e, n_planes, st);
    ctx->launches += launches + 1;
    CU(ctx, cudaStreamSynchronize(st));       // the host tables above are read by the asynchronous copies
    release();
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

int b2s_resize_aa(b2s_context *ctx, const void *d_in, int in_dtype, int rows, int cols, int new_rows, int new_cols,
                  const double *wy, int ry, const double *wx, int rx, float *d_out, int n_planes, void *stream)
{
    if (!ctx || !d_in || !d_out || rows <= 0 || cols <= 0 || new_rows <= 0 || new_cols <= 0 || n_planes <= 0 ||
        in_dtype < B2S_U8 || in_dtype > B2S_F32 || ry < 0 || rx < 0 || (ry && !wy) || (rx && !wx))
        return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (cudaMallocAsync(&p, bytes ? bytes : 16, st) != cudaSuccess) return nullptr;
        tmp.push_back(p);
        return p;
    };
    auto release = [&]() { for (void *p : tmp) cudaFreeAsync(p, st); tmp.clear(); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    const size_t elems = (size_t)rows * cols * n_planes;
    const int f64 = in_dtype != B2S_F32;          // skimage filters a float64 copy of an integer image
    const size_t esz = f64 ? 8 : 4;
    unsigned *mm = (unsigned *)dalloc(sizeof(unsigned) * 2 * n_planes);
    if (!mm) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
    CU(ctx, cudaMemsetAsync(mm, 0xff, sizeof(unsigned) * 2 * n_planes, st));
    b2s_launch_minmax(d_in, in_dtype, (size_t)rows * cols, n_planes, mm, st);
    int launches = 1;
    const void *cur = d_in;
    int cur_dt = in_dtype;
    const int radius[2] = {ry, rx};
    const double *w[2] = {wy, wx};
    for (int axis = 0; axis < 2; ++axis) {
        if (radius[axis] <= 0) continue;
        const int n = 2 * radius[axis] + 1;
        double *dw = (double *)dalloc(sizeof(double) * n);
        void *buf = dalloc(elems * esz);
        if (!dw || !buf) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
        CU(ctx, cudaMemcpyAsync(dw, w[axis], sizeof(double) * n, cudaMemcpyHostToDevice, st));
        b2s_launch_gauss_aa(cur, cur_dt, buf, f64, rows, cols, axis, dw, radius[axis], n_planes, st);
        cur = buf;
        cur_dt = f64 ? B2S_F64_INTERNAL : B2S_F32;
        ++launches;
    }
    std::vector<int> idx(2 * (size_t)(new_rows + new_cols));
    std::vector<double> wt(2 * (size_t)(new_rows + new_cols));
    b2s_resize_axis_table(rows, new_rows, idx.data(), idx.data() + new_rows, wt.data(), wt.data() + new_rows);
    b2s_resize_axis_table(cols, new_cols, idx.data() + 2 * new_rows, idx.data() + 2 * new_rows + new_cols,
                          wt.data() + 2 * new_rows, wt.data() + 2 * new_rows + new_cols);
    int *d_idx = (int *)dalloc(sizeof(int) * idx.size());
    double *d_w = (double *)dalloc(sizeof(double) * wt.size());
    if (!d_idx || !d_w) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
    CU(ctx, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_w, wt.data(), sizeof(double) * wt.size(), cudaMemcpyHostToDevice, st));
    B2sResizeArgs ra;
    ra.src = cur; ra.dtype = cur_dt; ra.mm_dtype = in_dtype; ra.rows = rows; ra.cols = cols;
    ra.new_rows = new_rows; ra.new_cols = new_cols;
    ra.iy0 = d_idx; ra.iy1 = ra.iy0 + new_rows; ra.ix0 = ra.iy1 + new_rows; ra.ix1 = ra.ix0 + new_cols;
    ra.wy0 = d_w; ra.wy1 = ra.wy0 + new_rows; ra.wx0 = ra.wy1 + new_rows; ra.wx1 = ra.wx0 + new_cols;
    ra.mm = mm;
    B2sEpilogueArgs e;
    memset(&e, 0, sizeof e);
    e.final_mode = 3; e.out_dtype = B2S_F32; e.out = d_out; e.out_rows = new_rows; e.out_cols = new_cols;
    b2s_launch_resize_final(ra, e, n_planes, st);
    ctx->launches += launches + 1;
    CU(ctx, cudaStreamSynchronize(st));       // the host tables above are read by the asynchronous copies
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

int b2s_isotropic_z(b2s_context *ctx, const float *d_in, int n_in, int64_t plane_elems, int method, float *d_out, void *stream)
{
    if (!ctx || !d_in || !d_out || n_in <= 0 || plane_elems <= 0 || method < B2S_DS_MAX || method > B2S_DS_MEAN) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    for (int k = 0; 2 * k < n_in; ++k) {
        const float *a = d_in + (size_t)(2 * k) * plane_elems;
        const float *b = 2 * k + 1 < n_in ? a + plane_elems : nullptr;
        b2s_launch_z_pair(a, b, method, d_out + (size_t)k * plane_elems, plane_elems, (cudaStream_t)stream);
        ctx->launches += 1;
    }
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

int b2s_isotropic_convert(b2s_context *ctx, const float *d_in, int64_t n, int mode, int shift, void *d_out, void *stream)
{
    if (!ctx || !d_in || !d_out || n <= 0 || (mode != 1 && mode != 2 && mode != 4)) return B2S_ERR_INVALID;
    if (mode == 2 && (shift < 0 || shift > 8)) return fail(ctx, B2S_ERR_INVALID, "right shift should be between 0 and 8");
    CU(ctx, cudaSetDevice(ctx->device));
    b2s_launch_convert_f32(d_in, n, mode, shift, d_out, (cudaStream_t)stream);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

int64_t b2s_deflate_bound(int dtype, int rows, int cols, int n_planes, int rows_per_strip)
{
    if (rows <= 0 || cols <= 0 || n_planes <= 0 || rows_per_strip <= 0 || dtype < 0 || dtype > 2) return B2S_ERR_INVALID;
    const size_t plane_bytes = (size_t)rows * cols * dtype_size(dtype);
    return (int64_t)b2s_deflate_bound_bytes(plane_bytes, (rows + rows_per_strip - 1) / rows_per_strip, n_planes);
}

int b2s_deflate_strips(b2s_context *ctx, const void *d_planes, int dtype, int rows, int cols, int n_planes, int rows_per_strip,
                       void *d_out, int64_t out_capacity, uint32_t *d_sizes, uint64_t *d_offsets, int64_t *total_bytes, void *stream)
{
    if (!ctx || !d_planes || !d_out || !d_sizes || !d_offsets || !total_bytes) return B2S_ERR_INVALID;
    if (rows <= 0 || cols <= 0 || n_planes <= 0 || rows_per_strip <= 0 || dtype < 0 || dtype > 2)
        return fail(ctx, B2S_ERR_INVALID, "b2s_deflate_strips: bad geometry");
    if (((uintptr_t)d_out & 3) != 0 || out_capacity < 64) return fail(ctx, B2S_ERR_INVALID, "b2s_deflate_strips: the output must be 4-byte aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t row_bytes = (size_t)cols * dtype_size(dtype), plane_bytes = row_bytes * rows;
    const int spp = (rows + rows_per_strip - 1) / rows_per_strip;
    const size_t n = (size_t)spp * n_planes;
    unsigned *tmp = nullptr;
    auto release = [&]() { if (tmp) cudaFreeAsync(tmp, st); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    if (cudaMallocAsync((void **)&tmp, sizeof(unsigned) * (n * 513 + 4), st) != cudaSuccess) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
    int *d_over = reinterpret_cast<int *>(tmp + n * 513);
    CU(ctx, cudaMemsetAsync(d_over, 0, sizeof(int), st));
    const size_t usable = ((size_t)out_capacity - 8) & ~(size_t)3;
    CU(ctx, cudaMemsetAsync(d_out, 0, (size_t)out_capacity, st));       // the bit writer ORs into shared boundary words
    b2s_launch_deflate(d_planes, plane_bytes, row_bytes, rows, rows_per_strip, n_planes, tmp, d_sizes,
                       reinterpret_cast<unsigned long long *>(d_offsets), d_out, usable, d_over, st);
    ctx->launches += 4;
    int over = 0;
    unsigned long long total = 0;
    CU(ctx, cudaMemcpyAsync(&over, d_over, sizeof over, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(&total, d_offsets + n, sizeof total, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    CU(ctx, cudaGetLastError());
    if (over) return fail(ctx, B2S_ERR_NOMEM, "b2s_deflate_strips: %llu compressed bytes do not fit the %lld-byte output (b2s_deflate_bound)",
                          total, (long long)out_capacity);
    *total_bytes = (int64_t)total;
    return B2S_OK;
}

int b2s_img_mask(b2s_context *ctx, const void *img, int dtype, int rows, int cols, int n_planes, double threshold, int close_steps,
                 int open_steps, unsigned char *mask, void *stream)
{
    if (!ctx || !img || !mask || rows <= 0 || cols <= 0 || n_planes <= 0) return B2S_ERR_INVALID;
    if (close_steps < 1 || open_steps < 1) return fail(ctx, B2S_ERR_INVALID, "b2s_img_mask: close_steps and open_steps must be at least 1");
    if (dtype != B2S_U8 && dtype != B2S_U16 && dtype != B2S_F32) return fail(ctx, B2S_ERR_INVALID, "b2s_img_mask: unknown dtype");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)rows * cols * n_planes;
    unsigned char *tmp = nullptr;
    int *h_flag = nullptr;
    auto release = [&]() { if (tmp) cudaFreeAsync(tmp, st); if (h_flag) cudaFreeHost(h_flag); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    if (cudaMallocAsync((void **)&tmp, 2 * n + 16, st) != cudaSuccess) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
    CU(ctx, cudaMallocHost((void **)&h_flag, sizeof(int)));
    if (dtype == B2S_F32) {
        B2sImg im;
        im.ptr = const_cast<float *>(reinterpret_cast<const float *>(img));
        im.plane_stride = (size_t)rows * cols; im.pitch = cols; im.rows = rows; im.cols = cols;
        b2s_launch_mask_threshold_f32(im, 0, rows, cols, threshold, nullptr, mask, n_planes, st);
    } else {
        b2s_launch_mask_threshold_int(img, dtype, (size_t)rows * cols, threshold, nullptr, mask, n_planes, st);
    }
    int *d_flag = reinterpret_cast<int *>(tmp + ((2 * n + 3) & ~(size_t)3));
    const int rc = b2s_launch_img_mask(mask, tmp, tmp + n, rows, cols, close_steps, open_steps, d_flag, h_flag, n_planes, st);
    ctx->launches += 12;
    if (rc) return fail(ctx, rc, "b2s_img_mask failed (rows of %d pixels)", cols);
    CU(ctx, cudaStreamSynchronize(st));
    return B2S_OK;
}

int b2s_histogram(b2s_context *ctx, const void *in, int in_is_device, int dtype, int64_t plane_elems, int n_planes,
                  uint64_t *hist, int hist_is_device, int per_plane, void *stream)
{
    if (!ctx || !in || !hist || plane_elems <= 0 || n_planes <= 0) return B2S_ERR_INVALID;
    if (dtype != B2S_U8 && dtype != B2S_U16) return fail(ctx, B2S_ERR_INVALID, "b2s_histogram takes uint8 or uint16 planes");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == B2S_U8 ? 1 : 2;
    const size_t n_hist = (size_t)65536 * (per_plane ? n_planes : 1);
    void *d_in = nullptr;
    unsigned long long *d_hist = reinterpret_cast<unsigned long long *>(hist);
    std::vector<void *> tmp;
    auto release = [&]() { for (void *p : tmp) cudaFreeAsync(p, st); tmp.clear(); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    if (!in_is_device) {
        // planes go through in slices so that a whole stack never needs a device copy of itself
        if (cudaMallocAsync(&d_in, esz * (size_t)plane_elems * std::min(n_planes, 16), st) != cudaSuccess)
            return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
        tmp.push_back(d_in);
    }
    if (!hist_is_device) {
        void *p = nullptr;
        if (cudaMallocAsync(&p, sizeof(uint64_t) * n_hist, st) != cudaSuccess) return fail(ctx, B2S_ERR_NOMEM, "cudaMallocAsync failed");
        tmp.push_back(p);
        d_hist = reinterpret_cast<unsigned long long *>(p);
        CU(ctx, cudaMemcpyAsync(d_hist, hist, sizeof(uint64_t) * n_hist, cudaMemcpyHostToDevice, st));   // counts are ADDED
    }
    if (in_is_device) {
        b2s_launch_histogram(in, dtype, (size_t)plane_elems, n_planes, d_hist, per_plane, st);
        ctx->launches += 1;
    } else {
        for (int z = 0; z < n_planes; z += 16) {
            const int nb = std::min(16, n_planes - z);
            CU(ctx, cudaMemcpyAsync(d_in, (const char *)in + (size_t)z * plane_elems * esz, esz * (size_t)plane_elems * nb,
                                    cudaMemcpyHostToDevice, st));
            b2s_launch_histogram(d_in, dtype, (size_t)plane_elems, nb, d_hist + (per_plane ? (size_t)z * 65536 : 0), per_plane, st);
            ctx->launches += 1;
        }
    }
    if (!hist_is_device) {
        CU(ctx, cudaMemcpyAsync(hist, d_hist, sizeof(uint64_t) * n_hist, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
    } else if (!in_is_device) {
        CU(ctx, cudaStreamSynchronize(st));   // the caller's host planes may be reused on return
    }
    CU(ctx, cudaGetLastError());
    return B2S_OK;
}

int b2s_is_uniform(b2s_context *ctx, const void *d_in, int dtype, int64_t n, int32_t *uniform, void *stream)
{
    if (!ctx || !d_in || !uniform || n <= 0 || dtype < B2S_U8 || dtype > B2S_F32) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *mm = nullptr, h[2];
    CU(ctx, cudaMallocAsync((void **)&mm, sizeof(unsigned) * 2, st));
    cudaMemsetAsync(mm, 0xff, sizeof(unsigned) * 2, st);
    b2s_launch_minmax(d_in, dtype, (size_t)n, 1, mm, st);
    ctx->launches += 1;
    cudaMemcpyAsync(h, mm, sizeof h, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(mm, st);
    CU(ctx, cudaStreamSynchronize(st));
    *uniform = h[0] == ~h[1];
    return B2S_OK;
}

int b2s_debug_math(b2s_context *ctx, int which, const float *in, float *out, int64_t n)
{
    if (!ctx || !in || !out || n < 0) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    float *d_in = nullptr, *d_out = nullptr;
    CU(ctx, cudaMalloc(&d_in, sizeof(float) * (n ? n : 1)));
    CU(ctx, cudaMalloc(&d_out, sizeof(float) * (n ? n : 1)));
    CU(ctx, cudaMemcpy(d_in, in, sizeof(float) * n, cudaMemcpyHostToDevice));
    if (n) b2s_launch_math(which, d_in, d_out, n, 0);
    ctx->launches += 1;
    CU(ctx, cudaMemcpy(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return B2S_OK;
}

int b2s_debug_expm1_table_check(b2s_context *ctx, int int_path, double dark, int work_dtype, uint64_t first, uint64_t count,
                                uint64_t *mismatches)
{
    if (!ctx || !mismatches || first + count > (1ull << 32)) return B2S_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    float *thr = nullptr;
    unsigned long long *d_bad = nullptr;
    CU(ctx, cudaMalloc(&thr, sizeof(float) * 65536 + sizeof(int)));
    CU(ctx, cudaMalloc(&d_bad, sizeof(unsigned long long)));
    CU(ctx, cudaMemset(d_bad, 0, sizeof(unsigned long long)));
    B2sEpiFn f;
    f.int_path = int_path; f.darkf = (float)dark; f.hi_w = work_dtype == B2S_U8 ? 255.f : 65535.f;
    int *d_kmax = reinterpret_cast<int *>(thr + 65536);
    b2s_launch_epi_thresholds(thr, f, d_kmax, 0);
    int kmax = 0;
    cudaError_t e = cudaMemcpy(&kmax, d_kmax, sizeof kmax, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
        b2s_launch_epi_table_check(thr, f, kmax, first, count, d_bad, 0);
        ctx->launches += 3;
        unsigned long long bad = 0;
        e = cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost);
        *mismatches = bad;
    }
    cudaFree(thr);
    cudaFree(d_bad);
    CU(ctx, e);
    return B2S_OK;
}

}  // extern "C"
