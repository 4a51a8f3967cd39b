/*
 * b2sio — native, multithreaded tile codec for the file boundary of the pystripe hot path.
 *
 * Replaces, for the formats the reference's Step 3 reads and writes, the per-file Python I/O of
 *   pystripe/core.py:200-264   imread_tif_raw_png  (tifffile / Pillow / raw memmap, one file per process)
 *   pystripe/core.py:276-334   imsave_tif          (tifffile.imwrite(compression=('ADOBE_DEFLATE', 1)))
 *   pystripe/raw.py:9-68       raw_imread / raw_imsave
 * with batch calls that decode many files at once straight into ONE caller-owned (page-locked) buffer — the batch the
 * GPU plan consumes — and encode a batch of result planes back to files, on a bounded number of host threads.
 *
 * Plain C ABI, no CUDA: this library only touches host memory.  Host binding: pystripe/_io.py (ctypes).
 * Supported TIFF subset (what light-sheet tile stacks are): classic and BigTIFF, II / MM byte order, one sample per
 * pixel, 8 / 16-bit unsigned or 32-bit float, strips or tiles, compression none (1), LZW (5), deflate (8, 32946), ZSTD (50000),
 * predictor none or horizontal differencing (2); the first IFD is the image.  PNG (core.py:209-210): greyscale, 8 or 16 bits,
 * not interlaced, every scanline filter.  Anything else returns B2SIO_ERR_UNSUPPORTED and the Python host falls back to
 * Pillow for that file, as the reference does.
 */
#ifndef B2SIO_H
#define B2SIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2SIO_VERSION 100

typedef enum b2sio_status {
    B2SIO_OK = 0,
    B2SIO_ERR_IO = -1,          /* open / read / write failed (errno in the message)                    */
    B2SIO_ERR_FORMAT = -2,      /* not a TIFF / .raw file, or a corrupt one                             */
    B2SIO_ERR_UNSUPPORTED = -3, /* valid file outside the subset above                                  */
    B2SIO_ERR_SHAPE = -4,       /* decoded shape / dtype differs from what the caller's buffer expects  */
    B2SIO_ERR_INVALID = -5
} b2sio_status;

/* dtype codes are those of b200stripe.h: 0 = uint8, 1 = uint16, 2 = float32 */
typedef struct b2sio_info {
    int32_t height, width;
    int32_t dtype;
    int32_t compression;   /* TIFF tag 259 (1, 5, 8, 32946); 0 for .raw                         */
    int32_t big_endian;    /* sample byte order in the file                                      */
    int32_t tiled;
    int64_t n_chunks;      /* strips or tiles                                                    */
} b2sio_info;

int b2sio_version(void);
/* message of the last failure on the calling thread */
const char *b2sio_last_error(void);

/* replaces: the header look-ups tifffile / raw_imread do before decoding (core.py:208, raw.py:17-40).
 * `path` ends in .tif / .tiff / .raw (case-insensitive); .raw uses the reference's endianness heuristic. */
int b2sio_probe(const char *path, b2sio_info *info);

/* replaces: imread_tif_raw_png for one file (core.py:200-264).  dst holds height x width samples of `dtype`
 * (native byte order on return); a file whose shape or dtype differs fails with B2SIO_ERR_SHAPE.
 * n_threads > 1 decodes the strips / tiles of this one file in parallel (whole stitched slices). */
int b2sio_read(const char *path, void *dst, int32_t height, int32_t width, int32_t dtype, int n_threads);

/* replaces: the per-file reads of one batch_filter batch (core.py:1515-1540 inside the worker farm, core.py:1687-1771).
 * File i is decoded into (char *)dst + i * plane_stride_bytes by one of n_threads workers; status[i] receives the
 * b2sio_status of file i (a failed file leaves its plane untouched; the others are still decoded).
 * Returns the number of files that failed. */
int b2sio_read_batch(const char *const *paths, int n_files, void *dst, size_t plane_stride_bytes, int32_t height,
                     int32_t width, int32_t dtype, int n_threads, int32_t *status);

/* replaces: imsave_tif (core.py:276-334): one classic little-endian TIFF, strips of whole rows.
 * deflate_level 0 = uncompressed (compression tag 1), 1..9 = ADOBE_DEFLATE (tag 8) at that zlib level, 100 + L (L = 1..22) =
 * Zstandard (tag 50000) at level L when libzstd can be loaded; strips compressed in parallel on n_threads.  The file is written under a temporary name and renamed, mode 0777 like the reference. */
int b2sio_write_tiff(const char *path, const void *src, int32_t height, int32_t width, int32_t dtype, int deflate_level,
                     int n_threads);

/* replaces: the per-file writes of one batch (core.py:1583-1593).  Plane i starts at (const char *)src +
 * i * plane_stride_bytes.  Returns the number of files that failed; status[i] per file. */
int b2sio_write_tiff_batch(const char *const *paths, int n_files, const void *src, size_t plane_stride_bytes,
                           int32_t height, int32_t width, int32_t dtype, int deflate_level, int n_threads, int32_t *status);

/* replaces: the same per-file writes when the strips were already deflated on the GPU (b2s_deflate_strips, include/b200stripe.h):
 * file i consists of strips_per_file streams, stream s at (const char *)data + strip_offsets[i * strips_per_file + s] with
 * strip_sizes[...] bytes; `compression` is the TIFF tag 259 value of the streams (8 = Adobe deflate).  paths[i] == NULL skips
 * slot i.  Returns the number of files that failed; status[i] per file. */
int b2sio_write_tiff_strips_batch(const char *const *paths, int n_files, const void *data, const uint64_t *strip_offsets,
                                  const uint32_t *strip_sizes, int32_t strips_per_file, int32_t rows_per_strip, int32_t height,
                                  int32_t width, int32_t dtype, int32_t compression, int n_threads, int32_t *status);

/* replaces: raw_imsave (raw.py:44-68): 8-byte header (width, height as native uint32) + uint16 samples */
int b2sio_write_raw(const char *path, const void *src, int32_t height, int32_t width);

#ifdef __cplusplus
}
#endif
#endif /* B2SIO_H */
