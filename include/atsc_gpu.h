/*
 * atsc_gpu.h -- C ABI of the B200-native ATSC hot path (libatsc_gpu.so).
 *
 * The reference (instaclustr/atsc v0.7.2, Rust) has no FFI seam; the seam this
 * library replaces is its in-process L1/L2 API (paths relative to atsc/src/):
 *
 *   Compressor::compress(&self, &[f64]) -> Vec<u8>                  compressor/mod.rs:63
 *   Compressor::compress_bounded(&self, &[f64], f64) -> Vec<u8>     compressor/mod.rs:76
 *   Compressor::get_compress_bounded_results(..) -> CompressorResult compressor/mod.rs:94
 *   Compressor::decompress(&self, usize, &[u8]) -> Vec<f64>         compressor/mod.rs:109
 *   CompressorFrame::compress_best(&mut self, &[f64], f32, usize)   frame/mod.rs:71
 *   CompressedStream::compress_chunk_bounded_with / decompress      data.rs:56 / data.rs:104
 *
 * Those are per-frame calls issued from a sequential loop (main.rs:146); a
 * per-frame FFI call would serialise the GPU, so the ABI is batch oriented:
 * many frames per call, caller-allocated outputs, integer status codes, never
 * unwinds.  A Rust `-sys` crate binds exactly these symbols.
 *
 * Plain pointers and sizes only.  `samples` / `out_samples` may be host or
 * device pointers (detected with cudaPointerGetAttributes); every other
 * pointer is host memory.
 */
#ifndef ATSC_GPU_H
#define ATSC_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compressor ids == bincode variant index of `enum Compressor`
 * (compressor/mod.rs:34-44); also what is serialised in each BRO frame. */
enum {
    ATSC_NOOP = 0,
    ATSC_FFT = 1,
    ATSC_IDW = 2,
    ATSC_CONSTANT = 3,
    ATSC_POLYNOMIAL = 4,
    ATSC_AUTO = 5,
    ATSC_RLE = 6
};

/* status codes (the reference panics instead; see SURVEY.md 8b "Errors") */
enum {
    ATSC_OK = 0,
    ATSC_ERR_ARG = 1,         /* bad argument (NULL, empty frame, speed > 6, ...) */
    ATSC_ERR_CUDA = 2,        /* CUDA runtime failure; see atsc_gpu_last_error */
    ATSC_ERR_CAPACITY = 3,    /* payload_buf too small; *payload_used = bytes needed */
    ATSC_ERR_UNSUPPORTED = 4, /* e.g. Auto passed to a non-bounded call (reference: todo!()) */
    ATSC_ERR_FORMAT = 5       /* malformed payload / BRO stream */
};

/* near_tie bits: a threshold in the reference's control flow was decided with a
 * margin smaller than the documented arithmetic tolerance (DESIGN.md "near-ties") */
enum {
    ATSC_TIE_FFT_LOOP = 1,  /* fft.rs:334 `(e*1000) as i32 < (err*1000) as i32` */
    ATSC_TIE_POLY_LOOP = 2, /* polynomial.rs:231/255 round3(e) vs round4(err) */
    ATSC_TIE_SELECT = 4,    /* frame/mod.rs:103,130,140 `error <= max_error` */
    ATSC_TIE_FFT_TOPK = 8   /* fft.rs:245-255 equal |z| at the top-k cut */
};

typedef struct atsc_ctx atsc_ctx;

/* One context owns streams + workspaces on each listed device.  Frames of a
 * call are sharded over the devices by contiguous range balanced on sample
 * count (frames are independent: no collective). device_ids==NULL -> device 0. */
int atsc_gpu_create(const int *device_ids, int n_devices, atsc_ctx **out);
void atsc_gpu_destroy(atsc_ctx *ctx);
const char *atsc_gpu_last_error(const atsc_ctx *ctx);

/* pinned host memory helpers (async H2D/D2H needs page-locked buffers) */
void *atsc_gpu_host_alloc(uint64_t bytes);
void atsc_gpu_host_free(void *p);

typedef struct {
    uint8_t compressor; /* chosen compressor (== requested unless Auto) */
    uint8_t near_tie;   /* ATSC_TIE_* bits */
    uint16_t iterations; /* refinement-loop iterations of the winning lossy compressor */
    uint32_t payload_len;
    uint64_t payload_off; /* into payload_buf */
    double error;         /* CompressorResult.error of the chosen candidate */
    /* diagnostics for Auto frames: candidates in the reference's order
     * [FFT, Polynomial, RLE] (frame/mod.rs:77); size 0 = not evaluated / pruned */
    double cand_error[3];
    uint32_t cand_size[3];
    uint32_t reserved;
} atsc_frame_out;

/*
 * Compress n_frames frames; frame i is samples[frame_off[i] .. +frame_len[i]].
 *
 *   bounded != 0  -> CompressedStream::compress_chunk_bounded_with (data.rs:56):
 *                    compressor==ATSC_AUTO runs compress_best (frame/mod.rs:71) with
 *                    `speed` in 0..6, anything else runs Compressor::compress_bounded
 *                    (compressor/mod.rs:76).
 *   bounded == 0  -> CompressedStream::compress_chunk_with (data.rs:47) ->
 *                    Compressor::compress (compressor/mod.rs:63); ATSC_AUTO is
 *                    ATSC_ERR_UNSUPPORTED (reference: todo!()).
 *   max_error     =  `E as f32 / 100.0` (main.rs:157); widened to f64 inside exactly
 *                    like frame/mod.rs:67,87.
 *
 * out[i] and the payload bytes (the frame's `data: Vec<u8>`) are written to host
 * memory.  Returns ATSC_ERR_CAPACITY with *payload_used = required bytes if
 * payload_cap is too small.
 */
int atsc_gpu_compress_frames(atsc_ctx *ctx, const double *samples, const uint64_t *frame_off,
                             const uint32_t *frame_len, uint32_t n_frames, uint8_t compressor,
                             float max_error, uint32_t speed, int bounded, atsc_frame_out *out,
                             uint8_t *payload_buf, uint64_t payload_cap, uint64_t *payload_used);

typedef struct {
    uint8_t compressor;   /* frame's compressor tag (0..6, not Auto) */
    uint32_t sample_count; /* CompressorFrame.sample_count */
    uint64_t payload_off; /* into payloads */
    uint32_t payload_len;
    uint64_t out_off; /* first output sample of this frame in out_samples */
} atsc_frame_in;

/* Compressor::decompress (compressor/mod.rs:109) for n_frames frames.  Every frame
 * writes exactly sample_count doubles at out_samples + out_off (Noop frames: the
 * stored vector length must equal sample_count, as the reference's writer guarantees). */
int atsc_gpu_decompress_frames(atsc_ctx *ctx, const atsc_frame_in *frames, uint32_t n_frames,
                               const uint8_t *payloads, uint64_t payload_bytes,
                               double *out_samples);

/* ---- stream level (host side of the path; mirrors main.rs:130-172) ---------- */

/* OptimizerPlan::clean_data + get_chunks_sizes (optimizer/mod.rs:64-98). */
uint64_t atsc_plan_chunk_sizes(uint64_t len, uint32_t *out_sizes, uint64_t cap);

/* compress_data (main.rs:130-166): clean, split, compress every frame on the GPU,
 * serialise header + bincode(Vec<CompressorFrame>) (data.rs:79-85, header.rs:60-67).
 * error_pct is the CLI's -e (0..50), speed the CLI's -c (0..6).
 * Series s is samples[series_off[s] .. +series_len[s]] (host memory); its .bro bytes
 * land at bro_buf + bro_off[s], length bro_len[s]. */
int atsc_gpu_compress_series(atsc_ctx *ctx, const double *samples, const uint64_t *series_off,
                             const uint64_t *series_len, uint32_t n_series, uint8_t compressor,
                             uint32_t error_pct, uint32_t speed, uint8_t *bro_buf,
                             uint64_t bro_cap, uint64_t *bro_off, uint64_t *bro_len,
                             uint8_t *frame_near_tie_any);

/* decompress_data (main.rs:168-172): parse each BRO stream (data.rs:89-103) and
 * expand all frames.  Pass out_samples==NULL to obtain sample counts only
 * (out_count[s]); otherwise series s is written at out_samples + out_off[s]. */
int atsc_gpu_decompress_series(atsc_ctx *ctx, const uint8_t *bro_buf, const uint64_t *bro_off,
                               const uint64_t *bro_len, uint32_t n_series, double *out_samples,
                               const uint64_t *out_off, uint64_t *out_count);

/* ---- host-only helpers either side of the path (no GPU needed) ------------------- */

/* contiguous frame ranges balanced by sample count: how a multi-device context (and bench.py's
 * ranks) shard a call.  out_first has n_parts + 1 entries; part p = frames [first[p], first[p+1]). */
void atsc_plan_shards(const uint32_t *frame_len, uint32_t n_frames, uint32_t n_parts, uint32_t *out_first);

/* WBRO container (wavbrro/src/wavbrro.rs:103-132): whole-file image <-> samples.
 * decode returns the sample count (copies min(count, cap)), -1 bad header, -2 corrupt archive;
 * encode returns the file size and writes it when it fits in cap. */
int64_t atsc_wbro_decode(const uint8_t *file, uint64_t len, double *out, uint64_t cap);
uint64_t atsc_wbro_encode(const double *samples, uint64_t n, uint8_t *out, uint64_t cap);

/* CSV value column (atsc/src/csv.rs:36-98).  has_header: locate time_field / value_field by name
 * (the time column is located but never parsed, like the reference); else first column.
 * Returns the value count, -1 time field missing, -2 value field missing, -3 parse failure. */
int64_t atsc_csv_read_values(const char *text, uint64_t len, int has_header, const char *time_field,
                             const char *value_field, double *out, uint64_t cap);

/* VSRI timestamp index (vsri/src/lib.rs): per series a list of line segments
 * [sample rate, first sample, first timestamp, samples]; text image "min\nmax\nm,x0,y0,n\n...".
 * The csv-compressor CLI stores it beside the .bro file (csv-compressor/src/main.rs:141-211).
 * Getters return 1 and the value, or 0 where the reference returns None. */
typedef struct atsc_vsri atsc_vsri;
atsc_vsri *atsc_vsri_new(void);
void atsc_vsri_free(atsc_vsri *v);
int32_t atsc_day_elapsed_seconds(int64_t timestamp_sec);           /* lib.rs:31-40 */
int atsc_vsri_update_for_point(atsc_vsri *v, int32_t y);           /* lib.rs:249-285; 1 = point in the past */
int32_t atsc_vsri_min(const atsc_vsri *v);
int32_t atsc_vsri_max(const atsc_vsri *v);
int32_t atsc_vsri_sample_count(const atsc_vsri *v);                /* lib.rs:368-371 */
uint64_t atsc_vsri_segment_count(const atsc_vsri *v);
int atsc_vsri_get_sample(const atsc_vsri *v, int32_t y, int32_t *x);          /* lib.rs:312-328 */
int atsc_vsri_get_time(const atsc_vsri *v, int32_t x, int32_t *y);            /* lib.rs:331-353 */
int atsc_vsri_get_next_sample(const atsc_vsri *v, int32_t y, int32_t *x);     /* lib.rs:156-172 */
int atsc_vsri_get_previous_sample(const atsc_vsri *v, int32_t y, int32_t *x); /* lib.rs:178-197 */
int atsc_vsri_is_empty(const atsc_vsri *v, int32_t t0, int32_t t1);           /* lib.rs:202-245 */
uint64_t atsc_vsri_all_timestamps(const atsc_vsri *v, int32_t *out, uint64_t cap); /* lib.rs:356-366 */
uint64_t atsc_vsri_to_text(const atsc_vsri *v, char *out, uint64_t cap);      /* lib.rs:442-462 */
atsc_vsri *atsc_vsri_from_text(const char *text, uint64_t len);               /* lib.rs:466-497; NULL = malformed */

/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t atsc_gpu_launch_count(const atsc_ctx *ctx);

/* device time (ms, CUDA events on the library's own streams) of the last compress / decompress call:
 * from the first operation issued to the last one completed, copies included; max over devices */
double atsc_gpu_last_call_ms(const atsc_ctx *ctx);

/* CUDA-event time (ms, summed over devices, accumulated since the last reset) of each kernel
 * on the library's own streams, 12 slots: [0] stats [1] plan+polynomial [2] rle [3] fft_fwd
 * [4] noop-size+select+scan [5] emit [6] decode [7] host time spent preparing and launching waves
 * [8] fft_small [9] fft (top-k + refinement loop) [10..11] reserved */
void atsc_gpu_kernel_ms(atsc_ctx *ctx, double *out12, int reset);

#ifdef __cplusplus
}
#endif
#endif /* ATSC_GPU_H */
