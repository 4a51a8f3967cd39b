// libb2sio.so — native multithreaded tile codec (include/b2sio.h): TIFF (none / LZW / deflate, strips or tiles, classic or
// BigTIFF, either byte order) and .raw tiles, decoded straight into one caller-owned batch buffer.
//
// Replaces the per-file Python I/O of the reference (pystripe/core.py:200-334, pystripe/raw.py:9-68); see the header for
// the entry-by-entry mapping.  Host-only code: g++ -O3, zlib for inflate / deflate, no CUDA.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2sio.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int dtype_size(int d) { return d == 0 ? 1 : (d == 1 ? 2 : 4); }

// Zstandard (TIFF compression 50000, the reference's other output option ('ZSTD', level)): libzstd is bound at run time —
// the image ships the library without its header; absent library => those files report B2SIO_ERR_UNSUPPORTED.
struct Zstd {
    size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
    size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
    size_t (*bound)(size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool ok = false;
    Zstd()
    {
        void *h = nullptr;
        for (const char *name : {"libzstd.so.1", "libzstd.so"})
            if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
        if (!h) return;
        decompress = (decltype(decompress))dlsym(h, "ZSTD_decompress");
        compress = (decltype(compress))dlsym(h, "ZSTD_compress");
        bound = (decltype(bound))dlsym(h, "ZSTD_compressBound");
        is_error = (decltype(is_error))dlsym(h, "ZSTD_isError");
        ok = decompress && compress && bound && is_error;
    }
};
const Zstd &zstd()
{
    static const Zstd z;
    return z;
}

bool host_is_le()
{
    const uint16_t v = 1;
    return *reinterpret_cast<const uint8_t *>(&v) == 1;
}

// run fn(i) for i in [0, n) on up to n_threads threads (the calling thread is one of them)
template <class F>
void parallel_for(int64_t n, int n_threads, F fn)
{
    if (n <= 0) return;
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
    if (T == 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int64_t> next{0};
    auto body = [&] { for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(body);
    body();
    for (auto &t : th) t.join();
}

struct Mapped {
    const uint8_t *p = nullptr;
    size_t size = 0;
    int fd = -1;
    int open_file(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(B2SIO_ERR_IO, "open(%s): %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) return fail(B2SIO_ERR_IO, "fstat(%s): %s", path, strerror(errno));
        size = (size_t)st.st_size;
        if (size == 0) return fail(B2SIO_ERR_FORMAT, "%s is empty", path);
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return fail(B2SIO_ERR_IO, "mmap(%s): %s", path, strerror(errno));
        madvise(m, size, MADV_SEQUENTIAL);
        p = (const uint8_t *)m;
        return 0;
    }
    // uncompressed samples go file -> destination with pread (no page faults on a mapping in between)
    bool pread_all(void *dst, size_t n, uint64_t off) const
    {
        uint8_t *q = (uint8_t *)dst;
        while (n) {
            const ssize_t k = ::pread(fd, q, n, (off_t)off);
            if (k < 0) { if (errno == EINTR) continue; return false; }
            if (k == 0) return false;
            q += k; off += (uint64_t)k; n -= (size_t)k;
        }
        return true;
    }
    ~Mapped()
    {
        if (p) munmap((void *)p, size);
        if (fd >= 0) ::close(fd);
    }
};

bool ends_with_ci(const char *s, const char *suffix)
{
    const size_t n = strlen(s), m = strlen(suffix);
    if (m > n) return false;
    for (size_t i = 0; i < m; ++i)
        if (tolower((unsigned char)s[n - m + i]) != suffix[i]) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------- TIFF directory
struct Tiff {
    bool le = true, big = false;
    int64_t width = 0, height = 0;
    int bits = 0, spp = 1, compression = 1, predictor = 1, sample_format = 1, planar = 1;
    int64_t rows_per_strip = 0, tile_w = 0, tile_h = 0;
    std::vector<uint64_t> offsets, counts;
    bool tiled = false;
    int dtype = -1;
};

struct Reader {
    const uint8_t *p;
    size_t size;
    bool le;
    bool ok(uint64_t off, uint64_t n) const { return off <= size && n <= size - off; }
    uint16_t u16(uint64_t o) const { return le ? (uint16_t)(p[o] | p[o + 1] << 8) : (uint16_t)(p[o] << 8 | p[o + 1]); }
    uint32_t u32(uint64_t o) const
    {
        return le ? (uint32_t)p[o] | (uint32_t)p[o + 1] << 8 | (uint32_t)p[o + 2] << 16 | (uint32_t)p[o + 3] << 24
                  : (uint32_t)p[o] << 24 | (uint32_t)p[o + 1] << 16 | (uint32_t)p[o + 2] << 8 | (uint32_t)p[o + 3];
    }
    uint64_t u64(uint64_t o) const { return le ? (uint64_t)u32(o) | (uint64_t)u32(o + 4) << 32 : (uint64_t)u32(o) << 32 | u32(o + 4); }
};

int type_size(int t)
{
    switch (t) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    default: return 0;
    }
}

int parse_tiff(const uint8_t *p, size_t size, const char *path, Tiff &t)
{
    if (size < 8) return fail(B2SIO_ERR_FORMAT, "%s: too short for a TIFF header", path);
    if (p[0] == 'I' && p[1] == 'I') t.le = true;
    else if (p[0] == 'M' && p[1] == 'M') t.le = false;
    else return fail(B2SIO_ERR_FORMAT, "%s: not a TIFF file (byte-order mark)", path);
    Reader r{p, size, t.le};
    const uint16_t magic = r.u16(2);
    uint64_t ifd;
    if (magic == 42) { t.big = false; ifd = r.u32(4); }
    else if (magic == 43) {
        if (size < 16 || r.u16(4) != 8) return fail(B2SIO_ERR_FORMAT, "%s: bad BigTIFF header", path);
        t.big = true;
        ifd = r.u64(8);
    } else return fail(B2SIO_ERR_FORMAT, "%s: not a TIFF file (magic %u)", path, magic);
    const int cnt_size = t.big ? 8 : 2, ent_size = t.big ? 20 : 12;
    if (!r.ok(ifd, cnt_size)) return fail(B2SIO_ERR_FORMAT, "%s: IFD offset beyond the file", path);
    const uint64_t n_ent = t.big ? r.u64(ifd) : r.u16(ifd);
    if (n_ent > 4096 || !r.ok(ifd + cnt_size, n_ent * ent_size)) return fail(B2SIO_ERR_FORMAT, "%s: corrupt IFD", path);
    std::vector<uint64_t> bits_v;
    for (uint64_t e = 0; e < n_ent; ++e) {
        const uint64_t o = ifd + cnt_size + e * ent_size;
        const int tag = r.u16(o), type = r.u16(o + 2);
        const uint64_t count = t.big ? r.u64(o + 4) : r.u32(o + 4);
        const uint64_t vo = o + (t.big ? 12 : 8);
        const int ts = type_size(type);
        if (!ts) continue;
        const uint64_t inline_cap = t.big ? 8 : 4;
        uint64_t data = vo;
        if (count * ts > inline_cap) data = t.big ? r.u64(vo) : r.u32(vo);
        auto fetch = [&](std::vector<uint64_t> &out) -> int {
            if (count > (1u << 26) || !r.ok(data, count * ts)) return fail(B2SIO_ERR_FORMAT, "%s: tag %d points beyond the file", path, tag);
            out.resize(count);
            for (uint64_t i = 0; i < count; ++i) {
                const uint64_t a = data + i * ts;
                out[i] = ts == 1 ? p[a] : ts == 2 ? r.u16(a) : ts == 4 ? r.u32(a) : r.u64(a);
            }
            return 0;
        };
        std::vector<uint64_t> v;
        switch (tag) {
        case 256: case 257: case 259: case 277: case 278: case 284: case 317: case 322: case 323: case 339: case 258: {
            if (int rc = fetch(v)) return rc;
            if (v.empty()) break;
            if (tag == 256) t.width = (int64_t)v[0];
            else if (tag == 257) t.height = (int64_t)v[0];
            else if (tag == 258) bits_v = v;
            else if (tag == 259) t.compression = (int)v[0];
            else if (tag == 277) t.spp = (int)v[0];
            else if (tag == 278) t.rows_per_strip = (int64_t)v[0];
            else if (tag == 284) t.planar = (int)v[0];
            else if (tag == 317) t.predictor = (int)v[0];
            else if (tag == 322) t.tile_w = (int64_t)v[0];
            else if (tag == 323) t.tile_h = (int64_t)v[0];
            else if (tag == 339) t.sample_format = (int)v[0];
            break;
        }
        case 273: case 324: if (int rc = fetch(t.offsets)) return rc; t.tiled = tag == 324; break;
        case 279: case 325: if (int rc = fetch(t.counts)) return rc; break;
        default: break;
        }
    }
    t.bits = bits_v.empty() ? 1 : (int)bits_v[0];
    if (t.width <= 0 || t.height <= 0 || t.width > (1 << 30) || t.height > (1 << 30))
        return fail(B2SIO_ERR_FORMAT, "%s: missing or absurd image size", path);
    if (t.spp != 1) return fail(B2SIO_ERR_UNSUPPORTED, "%s: %d samples per pixel (grey-scale tiles only)", path, t.spp);
    if (t.bits == 8 && t.sample_format == 1) t.dtype = 0;
    else if (t.bits == 16 && t.sample_format == 1) t.dtype = 1;
    else if (t.bits == 32 && t.sample_format == 3) t.dtype = 2;
    else return fail(B2SIO_ERR_UNSUPPORTED, "%s: %d-bit samples of format %d", path, t.bits, t.sample_format);
    if (t.compression != 1 && t.compression != 5 && t.compression != 8 && t.compression != 32946 &&
        !(t.compression == 50000 && zstd().ok))
        return fail(B2SIO_ERR_UNSUPPORTED, "%s: TIFF compression %d", path, t.compression);
    if (t.predictor != 1 && t.predictor != 2) return fail(B2SIO_ERR_UNSUPPORTED, "%s: TIFF predictor %d", path, t.predictor);
    if (t.predictor == 2 && t.dtype == 2) return fail(B2SIO_ERR_UNSUPPORTED, "%s: horizontal predictor on float samples", path);
    if (t.offsets.empty()) return fail(B2SIO_ERR_FORMAT, "%s: no strip / tile offsets", path);
    if (t.tiled) {
        if (t.tile_w <= 0 || t.tile_h <= 0) return fail(B2SIO_ERR_FORMAT, "%s: tiled without a tile size", path);
        const uint64_t need = (uint64_t)((t.width + t.tile_w - 1) / t.tile_w) * (uint64_t)((t.height + t.tile_h - 1) / t.tile_h);
        if (t.offsets.size() < need) return fail(B2SIO_ERR_FORMAT, "%s: %zu tiles, %llu needed", path, t.offsets.size(), (unsigned long long)need);
    } else {
        if (t.rows_per_strip <= 0 || t.rows_per_strip > t.height) t.rows_per_strip = t.height;
        const uint64_t need = (uint64_t)((t.height + t.rows_per_strip - 1) / t.rows_per_strip);
        if (t.offsets.size() < need) return fail(B2SIO_ERR_FORMAT, "%s: %zu strips, %llu needed", path, t.offsets.size(), (unsigned long long)need);
    }
    if (t.counts.size() < t.offsets.size()) {
        if (t.compression != 1) return fail(B2SIO_ERR_FORMAT, "%s: compressed data without byte counts", path);
        t.counts.assign(t.offsets.size(), 0);   // uncompressed: derived from the geometry below
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- decoders
// TIFF LZW (MSB-first codes, 9..12 bits, "early change"); returns the number of bytes produced
size_t lzw_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    static thread_local std::vector<uint16_t> prefix(4096);
    static thread_local std::vector<uint8_t> suffix(4096), first(4096);
    static thread_local std::vector<uint32_t> length(4096);
    for (int i = 0; i < 256; ++i) { suffix[i] = first[i] = (uint8_t)i; length[i] = 1; }
    uint64_t acc = 0;
    int have = 0, nbits = 9;
    size_t pos = 0, out = 0;
    int next = 258, old = -1;
    auto emit = [&](int code) -> bool {
        const uint32_t len = length[code];
        if (out + len > cap) {   // clip (a strip may carry padding)
            std::vector<uint8_t> tmp(len);
            int c = code;
            for (uint32_t k = len; k-- > 0;) { tmp[k] = suffix[c]; c = prefix[c]; }
            const size_t m = cap - out;
            memcpy(dst + out, tmp.data(), m);
            out = cap;
            return false;
        }
        int c = code;
        uint8_t *q = dst + out + len;
        for (uint32_t k = 0; k < len; ++k) { *--q = suffix[c]; c = prefix[c]; }
        out += len;
        return true;
    };
    for (;;) {
        while (have < nbits && pos < n) { acc = (acc << 8) | src[pos++]; have += 8; }
        if (have < nbits) break;
        const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1));
        have -= nbits;
        if (code == 257) break;
        if (code == 256) { nbits = 9; next = 258; old = -1; continue; }
        if (old < 0) {
            if (code >= 256) break;   // corrupt
            if (!emit(code)) break;
            old = code;
            continue;
        }
        if (code < next) {
            if (next < 4096) { prefix[next] = (uint16_t)old; suffix[next] = first[code]; first[next] = first[old]; length[next] = length[old] + 1; }
            if (!emit(code)) break;
        } else if (code == next && next < 4096) {
            prefix[next] = (uint16_t)old; suffix[next] = first[old]; first[next] = first[old]; length[next] = length[old] + 1;
            if (!emit(code)) break;
        } else break;   // corrupt
        if (next < 4096) ++next;
        if (next == (1 << nbits) - 1 && nbits < 12) ++nbits;
        old = code;
    }
    return out;
}

// decompress one strip / tile into `dst` (`want` bytes); short output is zero-filled
int decode_chunk(const Tiff &t, const uint8_t *src, size_t n, uint8_t *dst, size_t want, const char *path)
{
    if (t.compression == 1) {
        const size_t m = std::min(n ? n : want, want);
        memcpy(dst, src, m);
        if (m < want) memset(dst + m, 0, want - m);
        return 0;
    }
    if (t.compression == 5) {
        const size_t got = lzw_decode(src, n, dst, want);
        if (got < want) memset(dst + got, 0, want - got);
        return 0;
    }
    if (t.compression == 50000) {
        const size_t got = zstd().decompress(dst, want, src, n);
        if (zstd().is_error(got)) return fail(B2SIO_ERR_FORMAT, "%s: zstd decompression failed", path);
        if (got < want) memset(dst + got, 0, want - got);
        return 0;
    }
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) return fail(B2SIO_ERR_IO, "inflateInit failed");
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = (uInt)n;
    zs.next_out = dst;
    zs.avail_out = (uInt)want;
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = want - zs.avail_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) return fail(B2SIO_ERR_FORMAT, "%s: inflate failed (%d)", path, rc);
    if (got < want) memset(dst + got, 0, want - got);
    return 0;
}

// byte order, then the horizontal predictor, on `rows` rows of `cols` samples spaced `pitch_bytes`
void post_process(const Tiff &t, uint8_t *buf, int64_t rows, int64_t cols, size_t pitch_bytes)
{
    const int es = dtype_size(t.dtype);
    const bool swap = es > 1 && t.le != host_is_le();
    for (int64_t y = 0; y < rows; ++y) {
        uint8_t *row = buf + (size_t)y * pitch_bytes;
        if (swap) {
            if (es == 2) { uint16_t *q = (uint16_t *)row; for (int64_t x = 0; x < cols; ++x) q[x] = __builtin_bswap16(q[x]); }
            else { uint32_t *q = (uint32_t *)row; for (int64_t x = 0; x < cols; ++x) q[x] = __builtin_bswap32(q[x]); }
        }
        if (t.predictor == 2) {
            if (es == 1) { for (int64_t x = 1; x < cols; ++x) row[x] = (uint8_t)(row[x] + row[x - 1]); }
            else if (es == 2) { uint16_t *q = (uint16_t *)row; for (int64_t x = 1; x < cols; ++x) q[x] = (uint16_t)(q[x] + q[x - 1]); }
        }
    }
}

int read_tiff(const Mapped &m, const char *path, void *dst, int32_t height, int32_t width, int32_t dtype, int n_threads)
{
    Tiff t;
    if (int rc = parse_tiff(m.p, m.size, path, t)) return rc;
    if (t.height != height || t.width != width || t.dtype != dtype)
        return fail(B2SIO_ERR_SHAPE, "%s: %lld x %lld dtype %d, expected %d x %d dtype %d", path, (long long)t.height,
                    (long long)t.width, t.dtype, height, width, dtype);
    const int es = dtype_size(dtype);
    const size_t row_bytes = (size_t)width * es;
    uint8_t *out = (uint8_t *)dst;
    std::atomic<int> err{0};
    std::string first_err;
    std::atomic_flag err_lock = ATOMIC_FLAG_INIT;
    auto note = [&](int rc) {
        if (!rc) return;
        if (!err_lock.test_and_set()) { first_err = g_err; err.store(rc); }
    };
    if (!t.tiled && t.compression == 1) {
        // stored strips: row blocks of ~2 MB each, read by any thread (a single-strip file still decodes in parallel)
        struct Seg { int64_t y0, rows; uint64_t off; };
        std::vector<Seg> segs;
        const int64_t n_strips = (t.height + t.rows_per_strip - 1) / t.rows_per_strip;
        const int64_t seg_rows = std::max<int64_t>(1, (int64_t)((2u << 20) / row_bytes));
        for (int64_t s = 0; s < n_strips; ++s) {
            const int64_t y0 = s * t.rows_per_strip, rows = std::min(t.rows_per_strip, t.height - y0);
            for (int64_t r = 0; r < rows; r += seg_rows)
                segs.push_back({y0 + r, std::min(seg_rows, rows - r), t.offsets[s] + (uint64_t)r * row_bytes});
        }
        parallel_for((int64_t)segs.size(), n_threads, [&](int64_t i) {
            const Seg &g = segs[i];
            uint8_t *d = out + (size_t)g.y0 * row_bytes;
            const size_t want = (size_t)g.rows * row_bytes;
            size_t take = want;
            if (g.off >= m.size) take = 0;
            else if (want > m.size - g.off) take = (size_t)(m.size - g.off);   // truncated file: zero-filled
            if (take && !m.pread_all(d, take, g.off)) { note(fail(B2SIO_ERR_IO, "%s: read failed: %s", path, strerror(errno))); return; }
            if (take < want) memset(d + take, 0, want - take);
            post_process(t, d, g.rows, width, row_bytes);
        });
    } else if (!t.tiled) {
        const int64_t n_strips = (t.height + t.rows_per_strip - 1) / t.rows_per_strip;
        parallel_for(n_strips, n_threads, [&](int64_t s) {
            const int64_t y0 = s * t.rows_per_strip, rows = std::min(t.rows_per_strip, t.height - y0);
            const size_t want = (size_t)rows * row_bytes;
            const uint64_t off = t.offsets[s], cnt = t.counts[s];
            if (off > m.size || cnt > m.size - off) { note(fail(B2SIO_ERR_FORMAT, "%s: strip %lld beyond the file", path, (long long)s)); return; }
            uint8_t *d = out + (size_t)y0 * row_bytes;
            note(decode_chunk(t, m.p + off, (size_t)cnt, d, want, path));
            post_process(t, d, rows, width, row_bytes);
        });
    } else {
        const int64_t tx = (t.width + t.tile_w - 1) / t.tile_w, ty = (t.height + t.tile_h - 1) / t.tile_h;
        const size_t tile_row = (size_t)t.tile_w * es, tile_bytes = tile_row * (size_t)t.tile_h;
        parallel_for(tx * ty, n_threads, [&](int64_t i) {
            const int64_t iy = i / tx, ix = i - iy * tx;
            const uint64_t off = t.offsets[i], cnt = t.counts[i] ? t.counts[i] : tile_bytes;
            if (off > m.size || cnt > m.size - off) { note(fail(B2SIO_ERR_FORMAT, "%s: tile %lld beyond the file", path, (long long)i)); return; }
            std::vector<uint8_t> tmp(tile_bytes);
            note(decode_chunk(t, m.p + off, (size_t)cnt, tmp.data(), tile_bytes, path));
            post_process(t, tmp.data(), t.tile_h, t.tile_w, tile_row);
            const int64_t y0 = iy * t.tile_h, x0 = ix * t.tile_w;
            const int64_t rows = std::min(t.tile_h, t.height - y0), cols = std::min(t.tile_w, t.width - x0);
            for (int64_t y = 0; y < rows; ++y)
                memcpy(out + (size_t)(y0 + y) * row_bytes + (size_t)x0 * es, tmp.data() + (size_t)y * tile_row, (size_t)cols * es);
        });
    }
    if (err.load()) { g_err = first_err; return err.load(); }
    return 0;
}

// .raw: 8-byte header, reference heuristic for the byte order (raw.py:17-38)
int raw_header(const Mapped &m, const char *path, int64_t &h, int64_t &w, bool &be)
{
    if (m.size < 8) return fail(B2SIO_ERR_FORMAT, "%s: too short for a .raw header", path);
    Reader rl{m.p, m.size, true}, rb{m.p, m.size, false};
    const uint32_t w_le = rl.u32(0), h_le = rl.u32(4), w_be = rb.u32(0), h_be = rb.u32(4);
    if (w_le < w_be) { w = w_le; h = h_le; be = false; }
    else { w = w_be; h = h_be; be = true; }
    if (w <= 0 || h <= 0 || (uint64_t)w * (uint64_t)h * 2 > m.size - 8)
        return fail(B2SIO_ERR_FORMAT, "%s: header says %lld x %lld, file holds %zu bytes", path, (long long)h, (long long)w, m.size);
    return 0;
}

int read_raw(const Mapped &m, const char *path, void *dst, int32_t height, int32_t width, int32_t dtype, int n_threads)
{
    int64_t h, w;
    bool be;
    if (int rc = raw_header(m, path, h, w, be)) return rc;
    if (h != height || w != width || dtype != 1)
        return fail(B2SIO_ERR_SHAPE, "%s: %lld x %lld uint16, expected %d x %d dtype %d", path, (long long)h, (long long)w, height, width, dtype);
    const bool swap = be == host_is_le();
    const size_t row_bytes = (size_t)width * 2;
    const int64_t block = std::max<int64_t>(1, (1 << 20) / (int64_t)row_bytes);
    parallel_for((height + block - 1) / block, n_threads, [&](int64_t b) {
        const int64_t y0 = b * block, rows = std::min<int64_t>(block, height - y0);
        uint8_t *d = (uint8_t *)dst + (size_t)y0 * row_bytes;
        if (!m.pread_all(d, (size_t)rows * row_bytes, 8 + (uint64_t)y0 * row_bytes)) memset(d, 0, (size_t)rows * row_bytes);
        if (swap) { uint16_t *q = (uint16_t *)d; for (int64_t i = 0; i < rows * width; ++i) q[i] = __builtin_bswap16(q[i]); }
    });
    return 0;
}

// .png: greyscale, 8 or 16 bits, not interlaced (the tiles the reference reads through imageio's PNG-FI plugin,
// core.py:209-210).  One zlib stream over the IDAT chunks, scanline filters None / Sub / Up / Average / Paeth (PNG 1.2 §6),
// 16-bit samples big-endian.  Colour, palette, alpha or interlaced files are left to the caller's general reader.
struct Png { int64_t width = 0, height = 0; int depth = 0; };
int png_header(const Mapped &m, const char *path, Png &png)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (m.size < 33 || memcmp(m.p, sig, 8) != 0) return fail(B2SIO_ERR_FORMAT, "%s: not a PNG file", path);
    Reader r{m.p, m.size, false};
    if (r.u32(8) != 13 || memcmp(m.p + 12, "IHDR", 4) != 0) return fail(B2SIO_ERR_FORMAT, "%s: IHDR is not the first chunk", path);
    png.width = r.u32(16); png.height = r.u32(20); png.depth = m.p[24];
    const int colour = m.p[25], interlace = m.p[28];
    if (png.width <= 0 || png.height <= 0) return fail(B2SIO_ERR_FORMAT, "%s: empty image", path);
    if (colour != 0 || interlace != 0 || (png.depth != 8 && png.depth != 16))
        return fail(B2SIO_ERR_UNSUPPORTED, "%s: colour type %d, depth %d, interlace %d (greyscale 8 / 16 bit, not interlaced, is decoded here)",
                    path, colour, png.depth, interlace);
    return 0;
}

int read_png(const Mapped &m, const char *path, void *dst, int32_t height, int32_t width, int32_t dtype)
{
    Png png;
    if (int rc = png_header(m, path, png)) return rc;
    const int es = png.depth / 8;
    if (png.height != height || png.width != width || dtype != (es == 2 ? 1 : 0))
        return fail(B2SIO_ERR_SHAPE, "%s: %lld x %lld, %d bits, expected %d x %d dtype %d", path, (long long)png.height, (long long)png.width,
                    png.depth, height, width, dtype);
    const size_t row_bytes = (size_t)width * es, stride = row_bytes + 1;
    std::vector<uint8_t> raw(stride * (size_t)height);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) return fail(B2SIO_ERR_IO, "%s: inflateInit failed", path);
    zs.next_out = raw.data();
    zs.avail_out = (uInt)std::min<size_t>(raw.size(), 0xFFFFFFFFu);
    if (raw.size() > 0xFFFFFFFFull) { inflateEnd(&zs); return fail(B2SIO_ERR_UNSUPPORTED, "%s: image too large for this decoder", path); }
    Reader r{m.p, m.size, false};
    size_t pos = 8;
    int zrc = Z_OK;
    bool end = false;
    while (!end && pos + 12 <= m.size) {
        const uint32_t len = r.u32(pos);
        const uint8_t *type = m.p + pos + 4;
        if (pos + 12 + (size_t)len > m.size) { inflateEnd(&zs); return fail(B2SIO_ERR_FORMAT, "%s: truncated chunk", path); }
        if (memcmp(type, "IDAT", 4) == 0 && zrc == Z_OK) {
            zs.next_in = const_cast<Bytef *>(m.p + pos + 8);
            zs.avail_in = len;
            zrc = inflate(&zs, Z_NO_FLUSH);
            if (zrc != Z_OK && zrc != Z_STREAM_END) { inflateEnd(&zs); return fail(B2SIO_ERR_FORMAT, "%s: inflate failed (%d)", path, zrc); }
        } else if (memcmp(type, "IEND", 4) == 0) end = true;
        pos += 12 + (size_t)len;
    }
    const size_t got = raw.size() - zs.avail_out;
    inflateEnd(&zs);
    if (got != raw.size()) return fail(B2SIO_ERR_FORMAT, "%s: %zu of %zu bytes of scanlines", path, got, raw.size());
    // undo the scanline filters in place, then copy out (16 bit: big-endian -> native)
    const size_t bpp = (size_t)es;
    for (int64_t y = 0; y < height; ++y) {
        uint8_t *cur = raw.data() + (size_t)y * stride + 1;
        const uint8_t *up = y ? cur - stride : nullptr;
        const int ft = cur[-1];
        switch (ft) {
        case 0: break;
        case 1: for (size_t i = bpp; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
        case 2: if (up) for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
        case 3:
            for (size_t i = 0; i < row_bytes; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0;
                cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
            }
            break;
        case 4:
            for (size_t i = 0; i < row_bytes; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                cur[i] = (uint8_t)(cur[i] + (pa <= pb && pa <= pc ? a : (pb <= pc ? b : c)));
            }
            break;
        default: return fail(B2SIO_ERR_FORMAT, "%s: scanline filter %d", path, ft);
        }
        uint8_t *d = (uint8_t *)dst + (size_t)y * row_bytes;
        if (es == 1 || !host_is_le()) memcpy(d, cur, row_bytes);
        else for (int64_t x = 0; x < width; ++x) { d[2 * x] = cur[2 * x + 1]; d[2 * x + 1] = cur[2 * x]; }
    }
    return 0;
}

int read_any(const char *path, void *dst, int32_t height, int32_t width, int32_t dtype, int n_threads)
{
    if (!path || !dst || height <= 0 || width <= 0 || dtype < 0 || dtype > 2) return fail(B2SIO_ERR_INVALID, "bad argument");
    Mapped m;
    if (int rc = m.open_file(path)) return rc;
    if (ends_with_ci(path, ".raw")) return read_raw(m, path, dst, height, width, dtype, n_threads);
    if (ends_with_ci(path, ".png")) return read_png(m, path, dst, height, width, dtype);
    return read_tiff(m, path, dst, height, width, dtype, n_threads);
}

// ---------------------------------------------------------------------------------------------------- writer
struct Out {
    std::vector<uint8_t> b;
    void u16(uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
};

int write_all(int fd, const void *p, size_t n, const char *path)
{
    const uint8_t *q = (const uint8_t *)p;
    while (n) {
        const ssize_t k = ::write(fd, q, std::min<size_t>(n, (size_t)1 << 30));
        if (k < 0) { if (errno == EINTR) continue; return fail(B2SIO_ERR_IO, "write(%s): %s", path, strerror(errno)); }
        q += k;
        n -= (size_t)k;
    }
    return 0;
}

// layout: header | strip data | offsets array | counts array | IFD
int write_tiff_layout(const char *path, int32_t height, int32_t width, int32_t dtype, int64_t rps,
                      const std::vector<std::pair<const uint8_t *, size_t>> &pieces, int compression)
{
    const int es = dtype_size(dtype);
    const int64_t n_strips = (int64_t)pieces.size();
    if (n_strips == 0) return fail(B2SIO_ERR_INVALID, "%s: no strips to write", path);
    std::vector<uint32_t> offs(n_strips), cnts(n_strips);
    uint64_t pos = 8;
    for (int64_t s = 0; s < n_strips; ++s) {
        offs[s] = (uint32_t)pos;
        cnts[s] = (uint32_t)pieces[s].second;
        pos += pieces[s].second;
        if (pos > 0xFFF00000ull) return fail(B2SIO_ERR_UNSUPPORTED, "%s: file would exceed the classic TIFF limit", path);
    }
    if (pos & 1) ++pos;
    const uint32_t data_end = (uint32_t)pos;
    Out tail;
    uint32_t offs_at = 0, cnts_at = 0;
    if (n_strips > 1) {
        offs_at = data_end;
        for (uint32_t v : offs) tail.u32(v);
        cnts_at = data_end + (uint32_t)tail.b.size();
        for (uint32_t v : cnts) tail.u32(v);
    }
    const uint32_t ifd_at = data_end + (uint32_t)tail.b.size();
    struct Ent { uint16_t tag, type; uint32_t count, value; };
    std::vector<Ent> ents = {
        {256, 4, 1, (uint32_t)width}, {257, 4, 1, (uint32_t)height}, {258, 3, 1, (uint32_t)(8 * es)},
        {259, 3, 1, (uint32_t)compression}, {262, 3, 1, 1},
        {273, 4, (uint32_t)n_strips, n_strips > 1 ? offs_at : offs[0]}, {277, 3, 1, 1}, {278, 4, 1, (uint32_t)rps},
        {279, 4, (uint32_t)n_strips, n_strips > 1 ? cnts_at : cnts[0]}, {284, 3, 1, 1}, {339, 3, 1, (uint32_t)(dtype == 2 ? 3 : 1)},
    };
    tail.u16((uint16_t)ents.size());
    for (const Ent &e : ents) { tail.u16(e.tag); tail.u16(e.type); tail.u32(e.count); tail.u32(e.value); }
    tail.u32(0);
    Out head;
    head.b = {'I', 'I'};
    head.u16(42);
    head.u32(ifd_at);

    const std::string tmp = std::string(path) + ".b2s~";
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0777);
    if (fd < 0) return fail(B2SIO_ERR_IO, "open(%s): %s", tmp.c_str(), strerror(errno));
    int rc = write_all(fd, head.b.data(), head.b.size(), path);
    // consecutive pieces that are adjacent in memory go out in one write
    for (int64_t s = 0; s < n_strips && !rc;) {
        const uint8_t *p0 = pieces[s].first;
        size_t run = pieces[s].second;
        int64_t e = s + 1;
        while (e < n_strips && pieces[e].first == p0 + run) run += pieces[e++].second;
        rc = write_all(fd, p0, run, path);
        s = e;
    }
    if (!rc && data_end != offs.back() + cnts.back()) {   // the directory starts on a word boundary
        const uint8_t z = 0;
        rc = write_all(fd, &z, 1, path);
    }
    if (!rc) rc = write_all(fd, tail.b.data(), tail.b.size(), path);
    fchmod(fd, 0777);   // the reference chmods its output to 0o777 (core.py:311-314)
    if (::close(fd) != 0 && !rc) rc = fail(B2SIO_ERR_IO, "close(%s): %s", tmp.c_str(), strerror(errno));
    if (!rc && ::rename(tmp.c_str(), path) != 0) rc = fail(B2SIO_ERR_IO, "rename(%s): %s", path, strerror(errno));
    if (rc) ::unlink(tmp.c_str());
    return rc;
}

int write_tiff(const char *path, const void *src, int32_t height, int32_t width, int32_t dtype, int level, int n_threads)
{
    const bool use_zstd = level > 100;
    const int zlevel = use_zstd ? level - 100 : 0;
    if (!path || !src || height <= 0 || width <= 0 || dtype < 0 || dtype > 2 || level < 0 || (level > 9 && !(use_zstd && zlevel <= 22)))
        return fail(B2SIO_ERR_INVALID, "bad argument");
    if (use_zstd && !zstd().ok) return fail(B2SIO_ERR_UNSUPPORTED, "libzstd is not available");
    if (!host_is_le()) return fail(B2SIO_ERR_UNSUPPORTED, "the writer emits little-endian TIFF from a little-endian host only");
    const int es = dtype_size(dtype);
    const size_t row_bytes = (size_t)width * es, total = row_bytes * (size_t)height;
    if (total > 0xF0000000ull) return fail(B2SIO_ERR_UNSUPPORTED, "%s: %zu bytes need BigTIFF (not written by this codec)", path, total);
    const uint8_t *in = (const uint8_t *)src;
    // strips of whole rows: one strip when stored, ~256 KB of samples per strip when deflated (parallel + streamable)
    int64_t rps = height;
    if (level > 0) rps = std::max<int64_t>(1, std::min<int64_t>(height, (256 * 1024) / (int64_t)row_bytes));
    const int64_t n_strips = (height + rps - 1) / rps;
    std::vector<std::vector<uint8_t>> comp(level > 0 ? n_strips : 0);
    std::atomic<int> err{0};
    if (level > 0) {
        parallel_for(n_strips, n_threads, [&](int64_t s) {
            const int64_t y0 = s * rps, rows = std::min(rps, height - y0);
            const uLong n = (uLong)((size_t)rows * row_bytes);
            if (use_zstd) {
                comp[s].resize(zstd().bound(n));
                const size_t got = zstd().compress(comp[s].data(), comp[s].size(), in + (size_t)y0 * row_bytes, n, zlevel);
                if (zstd().is_error(got)) err.store(1);
                else comp[s].resize(got);
                return;
            }
            uLongf cap = compressBound(n);
            comp[s].resize(cap);
            if (compress2(comp[s].data(), &cap, in + (size_t)y0 * row_bytes, n, level) != Z_OK) err.store(1);
            comp[s].resize(cap);
        });
        if (err.load()) return fail(B2SIO_ERR_IO, "%s: deflate failed", path);
    }
    std::vector<std::pair<const uint8_t *, size_t>> pieces((size_t)n_strips);
    for (int64_t s = 0; s < n_strips; ++s) {
        const int64_t y0 = s * rps, rows = std::min(rps, height - y0);
        if (level > 0) pieces[s] = {comp[s].data(), comp[s].size()};
        else pieces[s] = {in + (size_t)y0 * row_bytes, (size_t)rows * row_bytes};
    }
    return write_tiff_layout(path, height, width, dtype, rps, pieces, use_zstd ? 50000 : (level > 0 ? 8 : 1));
}

}  // namespace

extern "C" {

int b2sio_version(void) { return B2SIO_VERSION; }
const char *b2sio_last_error(void) { return g_err.c_str(); }

int b2sio_probe(const char *path, b2sio_info *info)
{
    if (!path || !info) return fail(B2SIO_ERR_INVALID, "bad argument");
    memset(info, 0, sizeof *info);
    Mapped m;
    if (int rc = m.open_file(path)) return rc;
    if (ends_with_ci(path, ".raw")) {
        int64_t h, w;
        bool be;
        if (int rc = raw_header(m, path, h, w, be)) return rc;
        info->height = (int32_t)h; info->width = (int32_t)w; info->dtype = 1; info->big_endian = be; info->n_chunks = 1;
        return 0;
    }
    if (ends_with_ci(path, ".png")) {
        Png png;
        if (int rc = png_header(m, path, png)) return rc;
        info->height = (int32_t)png.height; info->width = (int32_t)png.width; info->dtype = png.depth == 16 ? 1 : 0;
        info->compression = 8; info->big_endian = 1; info->n_chunks = 1;
        return 0;
    }
    Tiff t;
    if (int rc = parse_tiff(m.p, m.size, path, t)) return rc;
    info->height = (int32_t)t.height; info->width = (int32_t)t.width; info->dtype = t.dtype;
    info->compression = t.compression; info->big_endian = !t.le; info->tiled = t.tiled;
    info->n_chunks = (int64_t)t.offsets.size();
    return 0;
}

int b2sio_read(const char *path, void *dst, int32_t height, int32_t width, int32_t dtype, int n_threads)
{
    return read_any(path, dst, height, width, dtype, n_threads);
}

int b2sio_read_batch(const char *const *paths, int n_files, void *dst, size_t plane_stride_bytes, int32_t height, int32_t width,
                     int32_t dtype, int n_threads, int32_t *status)
{
    if (!paths || n_files < 0 || !dst || !status) return fail(B2SIO_ERR_INVALID, "bad argument");
    std::atomic<int> failed{0};
    // many files: one worker per file; fewer files than workers: the spare workers split each file's strips
    const int per_file = n_files > 0 ? std::max(1, n_threads / std::max(1, n_files)) : 1;
    parallel_for(n_files, n_threads, [&](int64_t i) {
        status[i] = read_any(paths[i], (char *)dst + (size_t)i * plane_stride_bytes, height, width, dtype, per_file);
        if (status[i]) ++failed;
    });
    return failed.load();
}

int b2sio_write_tiff(const char *path, const void *src, int32_t height, int32_t width, int32_t dtype, int deflate_level, int n_threads)
{
    return write_tiff(path, src, height, width, dtype, deflate_level, n_threads);
}

int b2sio_write_tiff_batch(const char *const *paths, int n_files, const void *src, size_t plane_stride_bytes, int32_t height,
                           int32_t width, int32_t dtype, int deflate_level, int n_threads, int32_t *status)
{
    if (!paths || n_files < 0 || !src || !status) return fail(B2SIO_ERR_INVALID, "bad argument");
    std::atomic<int> failed{0};
    const int per_file = n_files > 0 ? std::max(1, n_threads / std::max(1, n_files)) : 1;
    parallel_for(n_files, n_threads, [&](int64_t i) {
        status[i] = write_tiff(paths[i], (const char *)src + (size_t)i * plane_stride_bytes, height, width, dtype, deflate_level, per_file);
        if (status[i]) ++failed;
    });
    return failed.load();
}

int b2sio_write_tiff_strips_batch(const char *const *paths, int n_files, const void *data, const uint64_t *strip_offsets,
                                  const uint32_t *strip_sizes, int32_t strips_per_file, int32_t rows_per_strip, int32_t height,
                                  int32_t width, int32_t dtype, int32_t compression, int n_threads, int32_t *status)
{
    if (!paths || n_files < 0 || !data || !strip_offsets || !strip_sizes || !status || strips_per_file <= 0 || rows_per_strip <= 0 ||
        height <= 0 || width <= 0 || dtype < 0 || dtype > 2 ||
        (int64_t)strips_per_file != ((int64_t)height + rows_per_strip - 1) / rows_per_strip)
        return fail(B2SIO_ERR_INVALID, "bad argument");
    std::atomic<int> failed{0};
    parallel_for(n_files, n_threads, [&](int64_t i) {
        status[i] = 0;
        if (!paths[i]) return;                     // a slot without a file (a plane that is not written)
        std::vector<std::pair<const uint8_t *, size_t>> pieces((size_t)strips_per_file);
        for (int32_t s = 0; s < strips_per_file; ++s) {
            const size_t k = (size_t)i * strips_per_file + s;
            pieces[s] = {(const uint8_t *)data + strip_offsets[k], (size_t)strip_sizes[k]};
        }
        status[i] = write_tiff_layout(paths[i], height, width, dtype, rows_per_strip, pieces, compression);
        if (status[i]) ++failed;
    });
    return failed.load();
}

int b2sio_write_raw(const char *path, const void *src, int32_t height, int32_t width)
{
    if (!path || !src || height <= 0 || width <= 0) return fail(B2SIO_ERR_INVALID, "bad argument");
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return fail(B2SIO_ERR_IO, "open(%s): %s", path, strerror(errno));
    const uint32_t head[2] = {(uint32_t)width, (uint32_t)height};
    int rc = write_all(fd, head, sizeof head, path);
    if (!rc) rc = write_all(fd, src, (size_t)height * width * 2, path);
    if (::close(fd) != 0 && !rc) rc = fail(B2SIO_ERR_IO, "close(%s): %s", path, strerror(errno));
    return rc;
}

}  // extern "C"
