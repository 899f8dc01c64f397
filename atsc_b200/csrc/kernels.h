// kernels.h -- host-callable launchers of the ATSC sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "fft.cuh"
#include "rle.cuh"
#include "stats.cuh"

namespace atsc {

// base pointers of per-CTA-slot workspaces; slot s uses base + s * stride
struct SlotPool {
    // RLE sort (rle.cuh)
    uint64_t *rle_k0, *rle_k1;
    uint32_t *rle_i0, *rle_i1, *rle_bnd;
    int rle_slots;
    // FFT (fft.cuh)
    float2 *fft_W, *fft_Xd, *fft_cD, *fft_cM;
    uint32_t *fft_keys, *fft_rank, *fft_locD, *fft_locM, *fft_ovr;
    FftEntry *fft_dlist;
    int fft_slots;
    int fwd_slots;  // CTAs of k_fft_fwd (they use fft_W only)
    // polynomial refinement loop (poly.cuh)
    double *poly_slope;  // [MAX_FRAME + 8] per slot
    int poly_slots;
    // decode scratch
    double *dec_pts;     // [MAX_FRAME + 8] per slot: decoded polynomial points / RLE values
    uint32_t *dec_mark;  // [MAX_FRAME + 8] per slot: RLE run-start markers
    uint32_t *dec_idx;   // [MAX_FRAME + 8] per slot
    int dec_slots;
};

struct DecFrame {
    uint64_t payload_off;
    uint64_t out_off;
    uint32_t payload_len;
    uint32_t sample_count;
    int32_t geom;
    uint8_t comp;
    uint8_t pad[3];
};

// compress pipeline; q = device array of >= 8 zeroed uint32 work-queue counters
void launch_stats(const FrameWork *fr, const ChunkRef *chunks, uint32_t n_chunks, const double *samples, StatsPart *parts,
                  unsigned *q, cudaStream_t st);
// p1_list / p1_count: k_poly1s' item descriptors are appended here (null: none)
void launch_plan(FrameWork *fr, uint32_t n, const double *samples, const StatsPart *parts, const FftGeom *geoms,
                 P1Item *p1_list, unsigned *p1_count, cudaStream_t st);
void launch_poly(FrameWork *fr, uint32_t n, const double *samples, double max_err, const double *inv_d2,
                 SlotPool pool, const double *first_parts, uint32_t parts_per_item, unsigned *q, cudaStream_t st);
// first candidate step of the frames with poly_parts != 0: items = (frame, part) pairs, one partial MAPE sum each
void launch_poly1(const FrameWork *fr, const ChunkRef *items, uint32_t n_items, const double *samples, double *parts,
                  unsigned *q, cudaStream_t st);
void launch_poly1s(const P1Item *list, const unsigned *count, uint32_t n_items, double *parts, cudaStream_t st);
void launch_rle(FrameWork *fr, uint32_t n, const double *samples, double max_err, SlotPool pool,
                unsigned *q, cudaStream_t st);
void launch_fft(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                SlotPool pool, FftEntry *arena, float2 *spec_xd, uint32_t *spec_keys, unsigned *q, cudaStream_t st);
void launch_fft_small(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                      FftEntry *arena, uint32_t lmax, unsigned *q, cudaStream_t st);
void launch_fft_fwd(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                    SlotPool pool, float2 *spec_xd, uint32_t *spec_keys, const float4 *fold_arena, unsigned *q,
                    cudaStream_t st);
// fused front end (front.cuh) of the frames items[0 .. n_items): frames with FM_ON set by the host
// probe tails (fft2.cuh: f2_probe_small) of the frames items[0 .. n_items) that k_sfold folded
void launch_probe(FrameWork *fr, const uint32_t *items, uint32_t n_items, double max_err, const FftGeom *geoms,
                  const float4 *fold_arena, unsigned *q, cudaStream_t st);
// stats + probe fold (sfold.cuh) of the frames with FM_SFOLD: items = (frame, first slot) pairs
void launch_sfold(const FrameWork *fr, const ChunkRef *items, uint32_t n_items, const double *samples, const FftGeom *geoms,
                  float4 *fold_arena, StatsPart *parts, unsigned *q, cudaStream_t st);
void launch_front(FrameWork *fr, const uint32_t *items, uint32_t n_items, const double *samples, double max_err,
                  const FftGeom *geoms, SlotPool pool, unsigned *q, cudaStream_t st);
constexpr uint32_t FRONT_MIN_SAMPLES = 16384;  // == FRONT_MIN_LEN (front.cuh)
void launch_noop_size(FrameWork *fr, uint32_t n, const double *samples, unsigned *q, cudaStream_t st);
void launch_select(FrameWork *fr, uint32_t n, double max_err, cudaStream_t st);
void launch_scan(FrameWork *fr, uint32_t n, unsigned long long *total, cudaStream_t st);
// writes nothing and raises *overflow when *total (k_scan's result) exceeds cap
void launch_emit(FrameWork *fr, uint32_t n, const double *samples, const FftGeom *geoms, SlotPool pool,
                 const FftEntry *arena, uint8_t *payload, const unsigned long long *total, unsigned long long cap,
                 unsigned *overflow, unsigned *q, cudaStream_t st);
// decompress
void launch_decode(const DecFrame *fr, uint32_t n, const uint8_t *payloads, double *out,
                   const FftGeom *geoms, SlotPool pool, const double *inv_d2, uint32_t *status,
                   unsigned *q, cudaStream_t st);
void launch_inv_d2(double *inv_d2, uint32_t n, cudaStream_t st);

int kernels_init();  // sets shared-memory attributes; returns cudaError_t as int

}  // namespace atsc
