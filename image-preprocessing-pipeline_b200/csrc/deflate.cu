// Deflate on the device: the result planes leave the GPU as the zlib streams of their TIFF strips.
//
// The reference writes every processed tile with tifffile's compression=('ADOBE_DEFLATE', 1) (pystripe/core.py:275-334,
// imsave_tif; batch_filter's default, core.py:1817).  On the host that deflate is the wall of the whole file path
// (≈1 GB/s on 16 threads) and the uncompressed result crosses PCIe first.  Here every strip of rows is compressed where it
// was computed, so D2H and the file write carry compressed bytes.
//
// Format: one zlib stream per strip (RFC 1950: 0x78 0x01, deflate data, Adler-32), one final dynamic-Huffman block
// (RFC 1951 §3.2.7) of literals only — no LZ77 matches.  On camera data almost all of zlib level 1's gain is the entropy
// coding of the bytes (measured on the reference's own tile: zlib-1 1.93x, this 1.92x), and a
// literal-only block has no serial dependency: a histogram, a code, a prefix sum of code lengths and a bit scatter.
// Block header: HLIT = 257, HDIST = 1 with a zero-length distance code ("no distance codes", §3.2.7), and a FIXED
// code-length code — the 16 lengths 0..15 get 4 bits each, 16 / 17 / 18 are unused — which is complete and costs
// 57 + 258 x 4 bits per strip (0.2 % of a 64 KB strip) instead of a second Huffman construction.
//
// Kernels: k_deflate_hist (CTA per strip, per-warp shared histograms), k_deflate_build (thread per strip: Huffman code
// lengths by the two-queue method on sorted leaves, frequencies halved until no code exceeds 15 bits, canonical codes
// stored bit-reversed, exact stream size), k_deflate_scan (offsets of the packed streams), k_deflate_encode (CTA per
// strip: per-thread chunk, block scan of chunk bit lengths, bit writer with atomicOr on the two boundary words of a chunk
// and plain stores in between, Adler-32 from per-chunk partial sums).  Pure integer / byte work; HBM traffic is one read
// of the plane per pass (hist, encode twice through L1) and one write of the compressed bytes.
#include <cstdint>
#include <cstdlib>

#include "b2s_internal.h"
#include "../../include/b200stripe.h"

namespace {

constexpr int kHdrBits = 16 + 3 + 5 + 5 + 4 + 19 * 3 + 258 * 4;   // zlib header + dynamic block header
constexpr int kNT = 256;

struct Geom {
    const unsigned char *in;
    size_t plane_bytes, row_bytes;
    int rows, rps, spp;          // rows per plane, rows per strip, strips per plane
};
__device__ __forceinline__ const unsigned char *strip_ptr(const Geom &g, int s, size_t &nbytes)
{
    const int plane = s / g.spp, k = s - plane * g.spp;
    const int r0 = k * g.rps, nr = min(g.rps, g.rows - r0);
    nbytes = (size_t)nr * g.row_bytes;
    return g.in + (size_t)plane * g.plane_bytes + (size_t)r0 * g.row_bytes;
}

__global__ void __launch_bounds__(kNT) k_deflate_hist(const Geom g, unsigned *hist)
{
    __shared__ unsigned h[kNT / 32][256];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kNT / 32) * 256; i += kNT) (&h[0][0])[i] = 0;
    __syncthreads();
    size_t n;
    const unsigned char *p = strip_ptr(g, blockIdx.x, n);
    unsigned *mine = h[warp];
    auto add = [&](unsigned b) {
        // camera data: the high bytes of a warp's samples are usually all equal — one add instead of a 32-way conflict
        const unsigned b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, b == b0)) { if ((threadIdx.x & 31) == 0) mine[b0] += 32; }
        else atomicAdd(&mine[b], 1u);
    };
    if (((uintptr_t)p & 3) == 0) {
        const unsigned *w = reinterpret_cast<const unsigned *>(p);
        const size_t nw = n >> 2, nw_full = nw / kNT * kNT;
        for (size_t i = threadIdx.x; i < nw_full; i += kNT) {        // whole warps: the uniform fast path applies
            const unsigned v = __ldg(w + i);
            add(v & 0xff); add((v >> 8) & 0xff); add((v >> 16) & 0xff); add(v >> 24);
        }
        for (size_t i = nw_full + threadIdx.x; i < nw; i += kNT) {
            const unsigned v = __ldg(w + i);
            atomicAdd(&mine[v & 0xff], 1u); atomicAdd(&mine[(v >> 8) & 0xff], 1u);
            atomicAdd(&mine[(v >> 16) & 0xff], 1u); atomicAdd(&mine[v >> 24], 1u);
        }
        for (size_t i = (nw << 2) + threadIdx.x; i < n; i += kNT) atomicAdd(&mine[p[i]], 1u);
    } else {
        for (size_t i = threadIdx.x; i < n; i += kNT) atomicAdd(&mine[p[i]], 1u);
    }
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int k = 0; k < kNT / 32; ++k) t += h[k][threadIdx.x];
    hist[(size_t)blockIdx.x * 256 + threadIdx.x] = t;
}

__device__ __forceinline__ unsigned bit_reverse(unsigned c, int len) { return __brev(c) >> (32 - len); }

// one thread per strip
__global__ void __launch_bounds__(32) k_deflate_build(const unsigned *hist, int n_strips, unsigned *tab, unsigned *sizes)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_strips) return;
    const unsigned *hs = hist + (size_t)s * 256;
    unsigned long long key[257];
    int n = 0;
    for (int sym = 0; sym < 256; ++sym) {
        const unsigned f = hs[sym];
        if (f) key[n++] = ((unsigned long long)f << 16) | (unsigned)sym;
    }
    key[n++] = (1ull << 16) | 256u;                       // end of block, once
    // ascending by (frequency, symbol): Shell sort, Ciura gaps
    const int gaps[6] = {132, 57, 23, 10, 4, 1};
    for (int gi = 0; gi < 6; ++gi) {
        const int gap = gaps[gi];
        for (int i = gap; i < n; ++i) {
            const unsigned long long v = key[i];
            int j = i;
            for (; j >= gap && key[j - gap] > v; j -= gap) key[j] = key[j - gap];
            key[j] = v;
        }
    }
    unsigned w[513];
    unsigned short par[513];
    unsigned char dep[513];
    for (;;) {
        for (int i = 0; i < n; ++i) w[i] = (unsigned)(key[i] >> 16);
        int li = 0, ii = n, next = n;
        while (next < 2 * n - 1) {                         // two queues: sorted leaves, internal nodes in creation order
            int a, b;
            if (li < n && (ii >= next || w[li] <= w[ii])) a = li++; else a = ii++;
            if (li < n && (ii >= next || w[li] <= w[ii])) b = li++; else b = ii++;
            w[next] = w[a] + w[b];
            par[a] = par[b] = (unsigned short)next;
            ++next;
        }
        int maxd = 0;
        if (n == 1) { dep[0] = 1; maxd = 1; }              // cannot happen (a strip has a byte and the end of block)
        else {
            dep[2 * n - 2] = 0;
            for (int id = 2 * n - 3; id >= 0; --id) {
                dep[id] = (unsigned char)(dep[par[id]] + 1);
                if (id < n) maxd = max(maxd, (int)dep[id]);
            }
        }
        if (maxd <= 15) break;
        for (int i = 0; i < n; ++i) {                      // flatten the distribution; the order stays sorted
            const unsigned long long f = ((key[i] >> 16) + 1) >> 1;
            key[i] = (f << 16) | (key[i] & 0xffffu);
        }
    }
    unsigned char len[257];
    for (int i = 0; i < 257; ++i) len[i] = 0;
    unsigned bl_count[16] = {0};
    for (int i = 0; i < n; ++i) {
        len[key[i] & 0xffffu] = dep[i];
        ++bl_count[dep[i]];
    }
    unsigned next_code[16];
    unsigned code = 0;
    bl_count[0] = 0;
    for (int bits = 1; bits <= 15; ++bits) {
        code = (code + bl_count[bits - 1]) << 1;
        next_code[bits] = code;
    }
    unsigned long long data_bits = 0;
    unsigned *ts = tab + (size_t)s * 257;
    for (int sym = 0; sym < 257; ++sym) {
        const int l = len[sym];
        if (l) {
            ts[sym] = ((unsigned)l << 16) | bit_reverse(next_code[l]++, l);
            data_bits += (unsigned long long)(sym < 256 ? hs[sym] : 1u) * l;
        } else ts[sym] = 0;
    }
    sizes[s] = (unsigned)((kHdrBits + data_bits + 7) / 8 + 4);
}

// ---- the same construction, one WARP per strip ------------------------------------------------------------------------
// The thread-per-strip kernel above keeps its arrays in local memory, where 32 threads walking 32 different arrays make
// every step a 32-sector access (0.77 ms for 4096 strips, 1.6 % occupancy).  Here the 256 literal keys are sorted by the warp
// in registers (bitonic, 8 keys per lane, as in lightsheet.cu), the arrays live in shared memory, and one lane runs the
// sequential parts (two queues, depths, canonical codes) at shared-memory latency while the table and the stream size are
// written by all lanes.  Same code lengths and codes as k_deflate_build by construction (same keys, same tie-breaks).
template <int K, int J>
__device__ __forceinline__ void bitonic_step(unsigned (&v)[8], int lane)
{
    if (J >= 8) {
        constexpr int M = J >> 3;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const unsigned other = __shfl_xor_sync(0xffffffffu, v[r], M);
            const bool up = (((lane << 3) | r) & K) == 0;
            const bool lower = (lane & M) == 0;
            v[r] = (lower == up) ? min(v[r], other) : max(v[r], other);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if ((r & J) == 0) {
                const unsigned x = v[r], y = v[r | J];
                const bool up = (((lane << 3) | r) & K) == 0;
                v[r] = up ? min(x, y) : max(x, y);
                v[r | J] = up ? max(x, y) : min(x, y);
            }
        }
    }
}
template <int K, int J>
struct BitonicMerge {
    static __device__ __forceinline__ void run(unsigned (&v)[8], int lane)
    {
        bitonic_step<K, J>(v, lane);
        BitonicMerge<K, (J >> 1)>::run(v, lane);
    }
};
template <int K>
struct BitonicMerge<K, 0> { static __device__ __forceinline__ void run(unsigned (&)[8], int) {} };
template <int K>
struct BitonicSort {
    static __device__ __forceinline__ void run(unsigned (&v)[8], int lane)
    {
        BitonicSort<(K >> 1)>::run(v, lane);
        BitonicMerge<K, (K >> 1)>::run(v, lane);
    }
};
template <>
struct BitonicSort<1> { static __device__ __forceinline__ void run(unsigned (&)[8], int) {} };

constexpr int kBuildWarps = 4;
__global__ void __launch_bounds__(kBuildWarps * 32) k_deflate_build_warp(const unsigned *hist, int n_strips, unsigned *tab, unsigned *sizes)
{
    __shared__ unsigned s_key[kBuildWarps][260];       // (frequency << 9) | symbol, ascending
    __shared__ unsigned s_w[kBuildWarps][516];
    __shared__ unsigned short s_par[kBuildWarps][516];
    __shared__ unsigned char s_dep[kBuildWarps][516];
    __shared__ unsigned char s_len[kBuildWarps][260];
    __shared__ unsigned short s_code[kBuildWarps][260];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kBuildWarps + warp;
    if (s >= n_strips) return;
    const unsigned *hs = hist + (size_t)s * 256;
    unsigned *key = s_key[warp], *w = s_w[warp];
    unsigned short *par = s_par[warp], *code = s_code[warp];
    unsigned char *dep = s_dep[warp], *len = s_len[warp];
    // element e = 8 * lane + r of the sort holds symbol e
    unsigned v[8];
    int used = 0, ones = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int sym = (lane << 3) | r;
        const unsigned f = hs[sym];
        v[r] = f ? ((f << 9) | (unsigned)sym) : 0xffffffffu;      // strips stay far below 2^23 bytes
        used += f != 0;
        ones += f == 1;
    }
    used = __reduce_add_sync(0xffffffffu, used);
    ones = __reduce_add_sync(0xffffffffu, ones);
    BitonicSort<256>::run(v, lane);
    // the end-of-block symbol (frequency 1, symbol 256) sorts right after the literals of frequency 1
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int e = (lane << 3) | r;
        if (e < used) key[e < ones ? e : e + 1] = v[r];
    }
    if (lane == 0) key[ones] = (1u << 9) | 256u;
    const int n = used + 1;
    for (int i = lane; i < 257; i += 32) len[i] = 0;
    __syncwarp();
    if (lane == 0) {
        for (;;) {
            for (int i = 0; i < n; ++i) w[i] = key[i] >> 9;
            int li = 0, ii = n, next = n;
            while (next < 2 * n - 1) {
                int a, b;
                if (li < n && (ii >= next || w[li] <= w[ii])) a = li++; else a = ii++;
                if (li < n && (ii >= next || w[li] <= w[ii])) b = li++; else b = ii++;
                w[next] = w[a] + w[b];
                par[a] = par[b] = (unsigned short)next;
                ++next;
            }
            int maxd = 0;
            if (n == 1) { dep[0] = 1; maxd = 1; }
            else {
                dep[2 * n - 2] = 0;
                for (int id = 2 * n - 3; id >= 0; --id) {
                    dep[id] = (unsigned char)(dep[par[id]] + 1);
                    if (id < n) maxd = max(maxd, (int)dep[id]);
                }
            }
            if (maxd <= 15) break;
            for (int i = 0; i < n; ++i) key[i] = ((((key[i] >> 9) + 1) >> 1) << 9) | (key[i] & 511u);
        }
        unsigned bl_count[16] = {0};
        for (int i = 0; i < n; ++i) {
            len[key[i] & 511u] = dep[i];
            ++bl_count[dep[i]];
        }
        unsigned next_code[16];
        unsigned c = 0;
        bl_count[0] = 0;
        for (int bits = 1; bits <= 15; ++bits) {
            c = (c + bl_count[bits - 1]) << 1;
            next_code[bits] = c;
        }
        for (int sym = 0; sym < 257; ++sym) {
            const int l = len[sym];
            code[sym] = l ? (unsigned short)bit_reverse(next_code[l]++, l) : 0;
        }
    }
    __syncwarp();
    unsigned long long bits = 0;
    unsigned *ts = tab + (size_t)s * 257;
    for (int sym = lane; sym < 257; sym += 32) {
        const unsigned l = len[sym];
        ts[sym] = l ? ((l << 16) | code[sym]) : 0u;
        bits += (unsigned long long)(sym < 256 ? hs[sym] : 1u) * l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if (lane == 0) sizes[s] = (unsigned)((kHdrBits + bits + 7) / 8 + 4);
}

// exclusive prefix of the strip sizes (bytes) -> offsets; total in offsets[n]
__global__ void __launch_bounds__(1024) k_deflate_scan(const unsigned *sizes, int n, unsigned long long *offsets)
{
    __shared__ unsigned long long part[1024];
    const int per = (n + 1023) / 1024, i0 = threadIdx.x * per, i1 = min(n, i0 + per);
    unsigned long long s = 0;
    for (int i = i0; i < i1; ++i) s += sizes[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { const unsigned long long t = part[i]; part[i] = run; run += t; }
        offsets[n] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (int i = i0; i < i1; ++i) { offsets[i] = run; run += sizes[i]; }
}

struct BitWriter {
    unsigned *base;
    unsigned long long acc;
    int nacc;
    size_t word;
    bool first;
    __device__ void init(unsigned *b, unsigned long long bitpos) { base = b; word = (size_t)(bitpos >> 5); nacc = (int)(bitpos & 31); acc = 0; first = true; }
    __device__ __forceinline__ void put(unsigned v, int n)          // n <= 24; bits are packed from the least significant end
    {
        acc |= (unsigned long long)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            if (first) { atomicOr(base + word, (unsigned)acc); first = false; }   // shared with the previous writer
            else base[word] = (unsigned)acc;
            acc >>= 32;
            nacc -= 32;
            ++word;
        }
    }
    __device__ void finish() { if (nacc > 0 && (unsigned)acc) atomicOr(base + word, (unsigned)acc); }
};

__global__ void __launch_bounds__(kNT) k_deflate_encode(const Geom g, const unsigned *tab, const unsigned long long *offsets,
                                                         unsigned *out, unsigned long long capacity, int *overflow)
{
    __shared__ unsigned t[257];
    __shared__ unsigned long long sc[kNT];
    __shared__ unsigned long long s1s[kNT], s2s[kNT];
    const int s = blockIdx.x;
    for (int i = threadIdx.x; i < 257; i += kNT) t[i] = tab[(size_t)s * 257 + i];
    size_t n;
    const unsigned char *p = strip_ptr(g, s, n);
    const unsigned long long byte0 = offsets[s], byte_end = offsets[s + 1];
    if (byte_end > capacity) { if (threadIdx.x == 0) *overflow = 1; return; }
    __syncthreads();
    size_t chunk = (n + kNT - 1) / kNT;
    chunk = (chunk + 3) & ~(size_t)3;
    const size_t c0 = min(n, threadIdx.x * chunk), c1 = min(n, c0 + chunk);
    const bool aligned = ((uintptr_t)p & 3) == 0;
    // pass A: bits, Adler partial sums of the chunk
    unsigned bits = 0;
    unsigned long long s1 = 0, s2 = 0;
    {
        const size_t L = c1 - c0;
        size_t i = c0;
        if (aligned) {
            for (; i + 4 <= c1; i += 4) {
                const unsigned v = __ldg(reinterpret_cast<const unsigned *>(p + i));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned b = (v >> (8 * k)) & 0xff;
                    bits += t[b] >> 16;
                    s1 += b;
                    s2 += (unsigned long long)(L - (i + k - c0)) * b;
                }
            }
        }
        for (; i < c1; ++i) {
            const unsigned b = p[i];
            bits += t[b] >> 16;
            s1 += b;
            s2 += (unsigned long long)(L - (i - c0)) * b;
        }
    }
    sc[threadIdx.x] = bits;
    s1s[threadIdx.x] = s1;
    s2s[threadIdx.x] = s2 + (unsigned long long)(n - c1) * s1;      // every byte after the chunk adds the chunk's sum once more
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0, a = 0, b = 0;
        for (int i = 0; i < kNT; ++i) {
            const unsigned long long v = sc[i];
            sc[i] = run;
            run += v;
            a += s1s[i];
            b += s2s[i] % 65521u;
        }
        s1s[0] = (1 + a) % 65521u;                                    // Adler-32 (RFC 1950 §8.2)
        s2s[0] = (b + n) % 65521u;
        s2s[1] = run;                                                 // data bits of the whole strip
    }
    __syncthreads();
    const unsigned long long bit_base = byte0 * 8;
    BitWriter wtr;
    if (threadIdx.x == 0) {
        wtr.init(out, bit_base);
        wtr.put(0x78, 8); wtr.put(0x01, 8);                           // zlib: deflate, 32 K window, no dictionary, fastest
        wtr.put(1, 1); wtr.put(2, 2);                                 // final block, dynamic Huffman codes
        wtr.put(0, 5); wtr.put(0, 5); wtr.put(15, 4);                 // 257 literal / length codes, 1 distance code, 19 code-length codes
        // code-length code lengths in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15: unused, then 4 bits for 0..15
        wtr.put(0, 9);
        for (int k = 0; k < 16; ++k) wtr.put(4, 3);
        for (int sym = 0; sym < 257; ++sym) wtr.put(__brev(t[sym] >> 16) >> 28, 4);   // canonical 4-bit code of a length = the length
        wtr.put(0, 4);                                                // the distance code: length 0
    } else {
        wtr.init(out, bit_base + kHdrBits + sc[threadIdx.x]);
    }
    {
        size_t i = c0;
        if (aligned) {
            for (; i + 4 <= c1; i += 4) {
                const unsigned v = __ldg(reinterpret_cast<const unsigned *>(p + i));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned e = t[(v >> (8 * k)) & 0xff];
                    wtr.put(e & 0xffffu, (int)(e >> 16));
                }
            }
        }
        for (; i < c1; ++i) {
            const unsigned e = t[p[i]];
            wtr.put(e & 0xffffu, (int)(e >> 16));
        }
    }
    if (threadIdx.x == kNT - 1) {
        // the last writer: end of block, zero bits up to the byte boundary, Adler-32 most significant byte first
        wtr.put(t[256] & 0xffffu, (int)(t[256] >> 16));
        const unsigned long long end_bits = kHdrBits + s2s[1] + (t[256] >> 16);
        const int pad = (int)((8 - (end_bits & 7)) & 7);
        if (pad) wtr.put(0, pad);
        const unsigned a = (unsigned)s1s[0], b = (unsigned)s2s[0];
        wtr.put(b >> 8, 8); wtr.put(b & 0xff, 8); wtr.put(a >> 8, 8); wtr.put(a & 0xff, 8);
    }
    wtr.finish();
}

}  // namespace

size_t b2s_deflate_bound_bytes(size_t plane_bytes, int strips_per_plane, int n_planes)
{
    // a Huffman code spends less than H + 1 <= 9.01 bits per byte; header, end of block, trailer and rounding per strip
    return (size_t)n_planes * (plane_bytes + plane_bytes / 8 + plane_bytes / 64 + (size_t)strips_per_plane * 160) + 64;
}

// tmp: hist (n_strips x 256 u32) | tab (n_strips x 257 u32); offsets: n_strips + 1 u64; out zeroed by the caller
void b2s_launch_deflate(const void *in, size_t plane_bytes, size_t row_bytes, int rows, int rows_per_strip, int n_planes, unsigned *tmp,
                        unsigned *sizes, unsigned long long *offsets, void *out, size_t capacity, int *overflow, cudaStream_t s)
{
    Geom g;
    g.in = (const unsigned char *)in; g.plane_bytes = plane_bytes; g.row_bytes = row_bytes; g.rows = rows; g.rps = rows_per_strip;
    g.spp = (rows + rows_per_strip - 1) / rows_per_strip;
    const int n = g.spp * n_planes;
    unsigned *hist = tmp, *tab = tmp + (size_t)n * 256;
    k_deflate_hist<<<n, kNT, 0, s>>>(g, hist);
    static const bool per_thread = getenv("B2S_DEFLATE_BUILD") && atoi(getenv("B2S_DEFLATE_BUILD")) == 0;
    if (per_thread || plane_bytes / g.spp >= (1u << 22)) k_deflate_build<<<(n + 31) / 32, 32, 0, s>>>(hist, n, tab, sizes);
    else k_deflate_build_warp<<<(n + kBuildWarps - 1) / kBuildWarps, kBuildWarps * 32, 0, s>>>(hist, n, tab, sizes);
    k_deflate_scan<<<1, 1024, 0, s>>>(sizes, n, offsets);
    k_deflate_encode<<<n, kNT, 0, s>>>(g, tab, offsets, (unsigned *)out, capacity, overflow);
}
