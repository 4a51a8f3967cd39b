// Stack statistics: exact intensity histograms of uint8 / uint16 planes.
//
// Replaces the sampling the reference does on the host before a channel is processed (process_images.py:594-659,
// `estimate_img_related_params`: log1p of three planes -> skimage.filters.threshold_multiotsu -> percentile above the top
// class -> bit shift / dark level / bleach clip levels).  Every statistic that routine needs is a function of the
// intensity histogram, and log1p is monotone on integers, so the GPU produces the exact 65 536-bin histogram of the integer
// pixels (any number of planes, at HBM speed) and the host maps bins through numpy's own log1p / histogram / percentile
// arithmetic (pystripe/stack_stats.py).  Histograms are additive: one process per GPU sums its planes, then ONE all-reduce
// of 65 536 int64 counters gives the whole-stack statistic — the only collective on this path.
//
// Kernel: a CTA owns a chunk of at most 65 535 pixels and counts them in shared memory with 16-bit counters packed two to
// a word (65 536 bins = 128 KB; a chunk cannot overflow a counter), then adds its non-zero words to the 64-bit global
// histogram.  Shared-memory atomics absorb the contention of the dominant background level; global atomics are one per
// occupied bin pair and chunk.
#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

constexpr int kHistThreads = 512;
constexpr int kChunk = 49152;   // pixels per CTA (< 65 536: the packed 16-bit counters cannot overflow)

template <typename T>
__global__ void __launch_bounds__(kHistThreads) k_hist(const T *in, size_t plane_elems, unsigned long long *hist, int per_plane)
{
    extern __shared__ unsigned s_cnt[];   // 32 768 words: bin v lives in half (v & 1) of word v >> 1
    constexpr int kWords = sizeof(T) == 1 ? 128 : 32768;
    for (int i = threadIdx.x; i < kWords; i += kHistThreads) s_cnt[i] = 0u;
    __syncthreads();
    const size_t plane = blockIdx.y;
    const size_t lo = (size_t)blockIdx.x * kChunk;
    const size_t hi = lo + kChunk < plane_elems ? lo + kChunk : plane_elems;
    const T *src = in + plane * plane_elems;
    for (size_t i = lo + threadIdx.x; i < hi; i += kHistThreads) {
        const unsigned v = src[i];
        atomicAdd(&s_cnt[v >> 1], 1u << (16 * (v & 1)));
    }
    __syncthreads();
    unsigned long long *out = hist + (per_plane ? plane * 65536 : 0);
    for (int i = threadIdx.x; i < kWords; i += kHistThreads) {
        const unsigned w = s_cnt[i];
        if (w & 0xffffu) atomicAdd(out + 2 * i, (unsigned long long)(w & 0xffffu));
        if (w >> 16) atomicAdd(out + 2 * i + 1, (unsigned long long)(w >> 16));
    }
}

}  // namespace

void b2s_launch_histogram(const void *in, int dtype, size_t plane_elems, int n_planes, unsigned long long *hist, int per_plane,
                          cudaStream_t s)
{
    const unsigned chunks = (unsigned)((plane_elems + kChunk - 1) / kChunk);
    if (dtype == B2S_U8) {
        k_hist<unsigned char><<<dim3(chunks, n_planes), kHistThreads, 128 * sizeof(unsigned), s>>>(
            reinterpret_cast<const unsigned char *>(in), plane_elems, hist, per_plane);
    } else {
        const int bytes = 32768 * sizeof(unsigned);
        cudaFuncSetAttribute(k_hist<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        k_hist<unsigned short><<<dim3(chunks, n_planes), kHistThreads, bytes, s>>>(
            reinterpret_cast<const unsigned short *>(in), plane_elems, hist, per_plane);
    }
}
