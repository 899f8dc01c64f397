// kernels.cu -- the ATSC hot path as sm_100a kernels.
//
// One CTA (1024 threads) works on one frame at a time and pulls frames from a device-side
// work queue, so a wave of frames of mixed sizes balances itself over the 148 SMs.
// Pipeline per wave (compress):  stats -> plan -> poly -> rle -> fft -> select -> scan -> emit
// Reference citations are relative to /root/reference/atsc/src/.
#include "kernels.h"

#include <cstdlib>

#include "fft2.cuh"
#include "fft_small.cuh"
#include "front.cuh"
#include "sfold.cuh"
#include "poly.cuh"
#include "stats.cuh"
#include "varscan.cuh"

namespace atsc {

__device__ inline RleWs rle_slot(const SlotPool &p, int s) {
    RleWs w;
    w.k0 = p.rle_k0 + (size_t)s * MAX_FRAME;
    w.k1 = p.rle_k1 + (size_t)s * MAX_FRAME;
    w.i0 = p.rle_i0 + (size_t)s * MAX_FRAME;
    w.i1 = p.rle_i1 + (size_t)s * MAX_FRAME;
    w.bnd = p.rle_bnd + (size_t)s * (MAX_FRAME + 8);
    return w;
}
__device__ inline FftWs fft_slot(const SlotPool &p, int s) {
    FftWs w;
    w.W = p.fft_W + (size_t)s * MAX_FFT_LEN;
    w.Xd = p.fft_Xd + (size_t)s * (MAX_FFT_LEN / 2 + 8);
    w.keys = p.fft_keys + (size_t)s * (MAX_FFT_LEN / 2 + 8);
    w.rank = p.fft_rank + (size_t)s * (MAX_FFT_LEN / 2 + 8);
    w.locD = p.fft_locD + (size_t)s * FFT_DEC_KCAP;
    w.locM = p.fft_locM + (size_t)s * FFT_DEC_KCAP;
    w.ovr = p.fft_ovr + (size_t)s * FFT_DEC_KCAP;
    w.cD = p.fft_cD + (size_t)s * FFT_DEC_KCAP;
    w.cM = p.fft_cM + (size_t)s * FFT_DEC_KCAP;
    w.dlist = p.fft_dlist + (size_t)s * FFT_DEC_KCAP;
    return w;
}

// =========================================================================================
// stats
// =========================================================================================
__global__ void __launch_bounds__(256, 8) k_stats(const FrameWork *__restrict__ fr, const ChunkRef *__restrict__ chunks,
                                                    uint32_t n_chunks, const double *__restrict__ samples,
                                                    StatsPart *parts, unsigned *q) {
    __shared__ StatsSmem sm;
    __shared__ int s_item;
    for (;;) {
        int c = queue_next(q, &s_item);
        if (c >= (int)n_chunks) break;
        const ChunkRef ch = chunks[c];
        const FrameWork *fw = &fr[ch.frame];
        chunk_stats(samples + fw->off, fw->len, ch.start, min(ch.start + STATS_CHUNK, fw->len), parts + c, &sm);
    }
}

// per frame: combine the chunk partials into the stats, then decide which candidates run
// (frame/mod.rs:71-149, compressor/mod.rs:63-107)
__global__ void k_plan(FrameWork *fr, uint32_t n, const double *__restrict__ samples, const StatsPart *__restrict__ parts,
                       const FftGeom *__restrict__ geoms, P1Item *__restrict__ p1_list, unsigned *p1_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameWork *fw = &fr[i];
    if (fw->front_mode & FM_ON) return;  // k_front finished this frame's stats and plan
    // partials: one per 32768-sample chunk (k_stats) or per slot-range item (k_sfold)
    const uint32_t nparts = (fw->front_mode & FM_SFOLD) ? sfold_items(geoms[fw->geom].M1) : (fw->len + STATS_CHUNK - 1) / STATS_CHUNK;
    finish_stats(samples + fw->off, fw->len, parts + fw->chunk0, nparts, fw);
    plan_frame(fw);
    // k_poly1s' work items: the frames poly_frame would evaluate the first step for (bounded Catmull-Rom, not
    // "same max and min"), cut into poly_parts balanced ranges of four-segment blocks (poly.cuh)
    if (p1_list && fw->poly_parts && fw->need_poly && !fw->poly_type && !fw->poly_valid && fw->bounded && fw->vmax != fw->vmin) {
        const uint32_t N = fw->len, Q = fw->poly_parts;
        const unsigned slot0 = atomicAdd(p1_count, Q);  // (poly_first_step(N) == 100: the host lists no other frame)
        for (uint32_t q = 0; q < Q; q++) {
            P1Item I;
            I.d = samples + fw->off;
            I.vmin = fw->vmin;
            I.vmax = fw->vmax;
            I.N = N;
            p1_item_blocks(N, q, &I.b_lo, &I.b_hi);
            I.out = fw->poly_part0 + q;
            I.nkeys = POLY_NS * (I.b_hi - I.b_lo) + 1;
            I.tame = poly_tame(fw->vmin, fw->vmax) ? 1u : 0u;
            I.pad[0] = I.pad[1] = I.pad[2] = I.pad[3] = 0u;
            p1_list[slot0 + q] = I;
        }
    }
}

// Fused front end of the big frames (front.cuh): stats + first Polynomial step + FFT probe fold in
// ONE pass over the samples, staged through shared memory by bulk asynchronous copies.
__global__ void __launch_bounds__(FR_CTA, 1) k_front(FrameWork *fr, const uint32_t *__restrict__ items, uint32_t n_items,
                                                     const double *__restrict__ samples, double max_err,
                                                     const FftGeom *__restrict__ geoms, SlotPool pool, unsigned *q,
                                                     uint32_t nap) {
    extern __shared__ __align__(128) unsigned char dyn_front[];
    FrontSmem *sm = reinterpret_cast<FrontSmem *>(dyn_front);
    uint32_t fill = 0, use = 0, issued = 0;
    if (threadIdx.x == FR_PRODUCER) {
        for (uint32_t s = 0; s < FR_SLOTS; s++) {
            mbar_init(&sm->full[s], 1u);
            mbar_init(&sm->empty[s], FR_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        front_claim(&sm->desc[0], fr, items, n_items, samples, geoms, q);
    }
    __syncthreads();
    for (uint32_t fc = 0;; fc++) {
        if (sm->desc[fc & 1u].idx >= n_items) break;  // the producer claims the next item while a frame streams
        front_frame(fr, items, n_items, fc, samples, max_err, geoms, pool.fft_W + (size_t)blockIdx.x * MAX_FFT_LEN, q, sm, fill,
                    use, issued, nap);
    }
}

// Stats + stage 1 of the FFT probe of the big Auto frames in one read (sfold.cuh)
__global__ void __launch_bounds__(SF_THREADS, SF_CTAS) k_sfold(const FrameWork *__restrict__ fr, const ChunkRef *__restrict__ items,
                                                               uint32_t n_items, const double *__restrict__ samples,
                                                               const FftGeom *__restrict__ geoms, float4 *__restrict__ fold_arena,
                                                               StatsPart *__restrict__ parts, unsigned *q) {
    __shared__ StatsSmem sm;
    __shared__ int s_item;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n_items) break;
        const ChunkRef ref = items[i];
        sfold_item(&fr[ref.frame], ref.start, samples, geoms, fold_arena, parts, &sm);
    }
}

// =========================================================================================
// polynomial / idw
// =========================================================================================
// first candidate step of the big Catmull-Rom frames, in balanced work items (poly.cuh: poly_first_step_item)
__global__ void __launch_bounds__(512, 2) k_poly1(const FrameWork *__restrict__ fr, const ChunkRef *__restrict__ items,
                                                 uint32_t n_items, const double *__restrict__ samples,
                                                 double *__restrict__ parts, unsigned *q) {
    __shared__ double shd[64];
    __shared__ double tang[POLY_ITEM_KEYS];
    __shared__ int s_item;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n_items) break;
        const ChunkRef ref = items[i];
        const FrameWork *fw = &fr[ref.frame];
        // the frames poly_frame would evaluate this step for (bounded Catmull-Rom, not "same max and min")
        if (!fw->need_poly || fw->poly_type || fw->poly_valid || !fw->bounded || fw->vmax == fw->vmin) continue;
        const double *d = samples + fw->off;
        double *out = parts + fw->poly_part0 + ref.start;
        if (poly_tame(fw->vmin, fw->vmax)) poly_first_step_item<true>(d, fw->len, fw->vmin, fw->vmax, ref.start, out, tang, shd);
        else poly_first_step_item<false>(d, fw->len, fw->vmin, fw->vmax, ref.start, out, tang, shd);
    }
}

// the same step from k_plan's compacted, self-contained item list with a static schedule (poly.cuh)
__global__ void __launch_bounds__(P1_T, 2) k_poly1s(const P1Item *__restrict__ items, const unsigned *__restrict__ count,
                                                   double *__restrict__ parts) {
    __shared__ P1Smem sm_;
    P1Smem *sm = &sm_;
    const uint32_t t = threadIdx.x, nb = gridDim.x, first = blockIdx.x;
    const uint32_t n_valid = *count;
    if (first >= n_valid) return;
    const uint32_t n_mine = (n_valid - first + nb - 1) / nb;  // entries first, first + nb, ...
    const uint32_t g = t / P1_STEP, j = t - g * P1_STEP;
    const bool active = g < P1_G;
    // raw keys j_lo - 1 .. j_hi + 1 of item kk into raw[kk & 1] (its descriptor is visible), one 8-byte copy per thread
    auto fetch_keys = [&](uint32_t kk) {
        const P1Item &X = sm->desc[kk & 3u];
        if (t < X.nkeys + 2u) {
            const uint32_t jj = POLY_NS * X.b_lo + t, kreg = (X.N + P1_STEP - 1u) / P1_STEP;  // key j_lo - 1 + t
            cp_async_8(&sm->raw[kk & 1u][t], X.d + (jj < kreg ? jj * P1_STEP : X.N - 1u));
        }
    };
    // ---- start-up: descriptors of the first two items, raw keys of the first, samples of the first trip
    if (t < 8u && (t >> 2) < n_mine)
        cp_async_16(reinterpret_cast<char *>(&sm->desc[t >> 2]) + (t & 3u) * 16u,
                    reinterpret_cast<const char *>(&items[first + (t >> 2) * nb]) + (t & 3u) * 16u);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    fetch_keys(0u);
    cp_async_commit();
    double o_nxt[POLY_NS] = {0.0, 0.0, 0.0, 0.0};
    if (active) p1_load_trip(o_nxt, sm->desc[0].d + (size_t)(1u + POLY_NS * (sm->desc[0].b_lo + g)) * P1_STEP + j);
    // Hermite basis of this thread's offset inside the segments (poly_cr_range)
    const double tt = __ddiv_rn((double)j, (double)P1_STEP);
    const double two_t = __dmul_rn(tt, 2.0), three_t = __dmul_rn(tt, 3.0);
    const double t2 = __dmul_rn(tt, tt), t3 = __dmul_rn(t2, tt);
    const double two_t3 = __dmul_rn(t2, two_t), two_t2 = __dmul_rn(tt, two_t), three_t2 = __dmul_rn(tt, three_t);
    const double h00 = __dadd_rn(__dsub_rn(two_t3, three_t2), 1.0), h10 = __dadd_rn(__dsub_rn(t3, two_t2), tt);
    const double h01 = __dsub_rn(three_t2, two_t3), h11 = __dsub_rn(t3, t2);
    for (uint32_t k = 0; k < n_mine; k++) {
        const P1Item &I = sm->desc[k & 3u];
        cp_async_wait<0>();  // this thread's copies of item k's raw keys / descriptor k+1 (issued one item ago, or above)
        __syncthreads();     // everybody is done with item k-1 (kt is free); desc[k], desc[k+1], raw[k & 1] are visible
        if (t < I.nkeys) {
            // tangent of key jj = j_lo + t (poly_mape's pre-pass): (v[jj+1] - v[jj-1]) / (pos[jj+1] - pos[jj-1]) * step
            const uint32_t jj = 1u + POLY_NS * I.b_lo + t, kreg = (I.N + P1_STEP - 1u) / P1_STEP;
            const uint32_t pa = (jj - 1u) * P1_STEP, pb = jj + 1u < kreg ? (jj + 1u) * P1_STEP : I.N - 1u;
            const double *r = sm->raw[k & 1u] + t;
            sm->kt[t] = make_double2(r[1], __dmul_rn(__ddiv_rn(__dsub_rn(r[2], r[0]), __dsub_rn((double)pb, (double)pa)), (double)P1_STEP));
        }
        __syncthreads();
        if (k + 2u < n_mine && t < 4u)
            cp_async_16(reinterpret_cast<char *>(&sm->desc[(k + 2u) & 3u]) + t * 16u,
                        reinterpret_cast<const char *>(&items[first + (k + 2u) * nb]) + t * 16u);
        if (k + 1u < n_mine) fetch_keys(k + 1u);
        cp_async_commit();
        const double2 *kp = sm->kt + POLY_NS * g;  // key 1 + NS * (b_lo + g) of the item's first trip
        double acc = I.tame ? p1_item_loop<true>(I, kp, o_nxt, active, g, j, h00, h10, h01, h11)
                            : p1_item_loop<false>(I, kp, o_nxt, active, g, j, h00, h10, h01, h11);
        if (k + 1u < n_mine && active) {  // the next item's first trip, in flight across the reduction and the barriers
            const P1Item &Nx = sm->desc[(k + 1u) & 3u];
            p1_load_trip(o_nxt, Nx.d + (size_t)(1u + POLY_NS * (Nx.b_lo + g)) * P1_STEP + j);
        }
        acc = warp_sum(acc);
        if ((t & 31u) == 0u) parts[(size_t)I.out * P1_PARTS + (t >> 5)] = acc;
    }
}

__global__ void __launch_bounds__(512, 2) k_poly(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                                double max_err, const double *__restrict__ inv_d2, SlotPool pool,
                                                const double *__restrict__ first_parts, uint32_t parts_per_item, unsigned *q) {
    __shared__ double shd[64];
    __shared__ int s_item;
    PolyWs ws;
    ws.slope = pool.poly_slope + (size_t)blockIdx.x * (MAX_FRAME + 8);
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (!fw->need_poly || fw->poly_valid) continue;  // poly_valid: k_front settled it at the first step
        poly_frame(samples + fw->off, fw, max_err, inv_d2, shd, ws, first_parts, parts_per_item);
    }
}

// =========================================================================================
// rle
// =========================================================================================
__global__ void __launch_bounds__(BLOCK) k_rle(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                               double max_err, SlotPool pool, unsigned *q) {
    extern __shared__ uint32_t dyn_hist[];
    __shared__ double shd[64];
    __shared__ int s_item;
    RleWs ws = rle_slot(pool, blockIdx.x);
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (!fw->need_rle) continue;
        if (fw->comp == C_AUTO && fw->forced == 0xFF) {
            // RLE comes last in the candidate list (frame/mod.rs:77) and min_by_key keeps the first
            // minimum, so it can only win with a strictly smaller payload than every passing
            // candidate before it.  Polynomial and FFT have already run: when even a lower bound of
            // the RLE size loses, the sort is skipped.
            uint32_t best = 0xFFFFFFFFu;
            if (fw->poly_valid == 1 && fw->poly_err <= max_err) best = min(best, fw->poly_size);
            if (fw->fft_valid == 1 && fw->fft_err <= max_err) best = min(best, fw->fft_size);
            uint32_t lb = rle_lower_bound(fw);
            if (lb >= best) {
                if (threadIdx.x == 0) {
                    fw->rle_valid = 2;
                    fw->rle_size = lb;
                }
                continue;
            }
        }
        rle_process(samples + fw->off, fw, ws, nullptr, (uint32_t *)shd, dyn_hist);
    }
}

// =========================================================================================
// fft
// =========================================================================================
// k_fft: one 512-thread CTA per SM (128 registers for the radix-27 butterflies of f2_inverse)
constexpr int K_FFT_SMEM_BYTES = F2I_SMEM_F2 * (int)sizeof(float2) > FFT_SMEM_BYTES ? F2I_SMEM_F2 * (int)sizeof(float2) : FFT_SMEM_BYTES;
__device__ inline bool fft_loop_near_tie(double cur, int E) {
    if (!(cur == cur) || isinf(cur)) return false;
    double c = cur * 1000.0;
    double nearest = round(c);
    return fabs(c - nearest) < 2e-4 && nearest == (double)E + 1.0;
}

__device__ void fft_frame(const double *__restrict__ d, FrameWork *fw, const FftGeom *__restrict__ geoms,
                          FftWs ws, FftEntry *list, double max_err, float2 *sm, double *shd, FftGeom *sg,
                          float2 *spec_xd, uint32_t *spec_keys) {
    uint32_t *sh = (uint32_t *)shd;
    const uint32_t N = fw->len;
    const bool bounded = fw->bounded != 0;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    if (fw->f32_const) {
        // fft.rs:289-292 "Same max and min": no frequencies, error None -> 0.0
        if (t == 0) {
            fw->fft_count = 0;
            fw->fft_err = 0.0;
            fw->fft_size = fft_payload_size(0, 0);
            fw->fft_iters = 0;
            fw->fft_tie = 0;
            fw->fft_valid = 1;
        }
        return;
    }
    const float vminf = (float)fw->vmin, vmaxf = (float)fw->vmax;
    const uint32_t mf = (3 >= N / 100) ? 3 : N / 100;
    const uint32_t hstep = max(mf / 2, 1u), tstep = max(mf / 10, 1u);
    const uint32_t kmax = bounded ? mf + 17 * hstep + 5 * tstep : mf;
    const int gi = fw->geom;
    uint32_t L, Bn, prefix = 0;
    __syncthreads();
    if (gi >= 0) {
        if (t == 0) *sg = geoms[gi];
        __syncthreads();
        L = sg->L;
        Bn = sg->Bn;
        if (bounded && N >= 128) prefix = (L - N) / 2;
        if (fw->fwd_done) {
            // k_fft_fwd (fft2.cuh) already produced the half spectrum and its keys
            ws.Xd = spec_xd + fw->spec_off;
            ws.keys = spec_keys + fw->spec_off;
        } else {
            fft_forward(d, N, prefix, *sg, ws, sm);
        }
    } else {
        // direct DFT, no padding
        L = N;
        Bn = N / 2 + 1;
        for (uint32_t k = t; k < Bn; k += T) {
            float2 acc = make_float2(0.f, 0.f);
            for (uint32_t j = 0; j < N; j++) {
                float2 w = unit_root(k * j, N, false);
                float x = (float)d[j];
                acc.x += x * w.x;
                acc.y += x * w.y;
            }
            ws.Xd[k] = acc;
            double nr = sqrt((double)acc.x * (double)acc.x + (double)acc.y * (double)acc.y);
            ws.keys[k] = __float_as_uint((float)nr);
        }
        __syncthreads();
    }
    const bool alias = Bn > 65536u;
    // a later candidate can only lose to FFT on size; FFT wins ties (frame/mod.rs:77,104,141)
    uint32_t bound = 0xFFFFFFFFu;
    if (bounded && fw->comp == C_AUTO && fw->forced == 0xFF) {
        if (fw->poly_valid == 1 && fw->poly_err <= max_err) bound = min(bound, fw->poly_size);
        if (fw->need_rle) bound = min(bound, fw->rle_valid == 1 ? fw->rle_size : rle_upper_bound(fw));  // RLE's error is 0: it always passes
    }
    if (bound != 0xFFFFFFFFu && !fw->fwd_done) {
        // Early exit before any sorting: the first schedule point keeps c1 = min(max_freq, #nonzero
        // bins) entries (fft.rs:249-252 stops at an exact zero) and the payload only grows from
        // there.  At most `smax` of them can have a one-byte position (pos < 251 after the u16 wrap).
        const uint32_t nz = fft_count_nonzero(Bn, ws, sh);
        const uint32_t c1 = min(min(mf, nz), min(fw->fft_list_cap, (uint32_t)FFT_KCAP));
        const uint32_t smax = alias ? 502u : 251u;
        if (fft_payload_size(c1, min(c1, smax)) > bound) {
            if (t == 0) {
                fw->fft_count = c1;
                fw->fft_err = max_err + 1.0;
                fw->fft_size = 0;
                fw->fft_iters = 1;
                fw->fft_tie = 0;
                fw->fft_valid = 2;
            }
            return;
        }
    }
    bool tie_cut;
    unsigned long long *S = (unsigned long long *)sm;
    const uint32_t pM = (gi >= 0 && sg->real) ? sg->M : 0u, pM1 = gi >= 0 ? sg->M1 : 1u, pM2 = gi >= 0 ? sg->M2 : 1u;
    const uint32_t kcap = min(kmax, fw->fft_list_cap);
    uint32_t K = 0, cutmask = 0;
    bool list_full = false;
    // Builds the sorted list of the `want` largest bins (a prefix of the full list: the order
    // is deterministic), the per-entry scatter coefficients, and the mask of schedule cuts that
    // split equal |z| (BinaryHeap pop order among equals is unspecified -> near-tie flag).
    auto build_list = [&](uint32_t want) {
        K = fft_topk(Bn, ws, want, list, S, sh, &tie_cut, pM, pM1, pM2);
        list_full = (want >= kcap) || (K < want);
        if (t == 0) {
            uint32_t mask = 0, jump = 0;
            for (int it = 1; it <= FFT_SCHED; it++) {
                uint32_t c = min(mf + jump, K);
                if (fft_cut_splits_tie_list(list, c, K) || (c == K && tie_cut && mf + jump == K)) mask |= 1u << (it - 1);
                jump += it <= 17 ? hstep : tstep;
                if (!bounded || mf + jump > K) break;
            }
            sh[105] = mask;
        }
        __syncthreads();
        cutmask |= sh[105];
        __syncthreads();
        if (bounded && gi >= 0) fft_prepare_entries(*sg, ws, list, K, alias);
    };
    // most frames stop after one or two refinement iterations: sort only the first two schedule
    // points' worth of bins up front, the full list (<= 16384 entries) only when needed
    build_list(bounded ? min(mf + hstep, kcap) : kcap);

    auto nsmall = [&](uint32_t c) -> uint32_t {
        uint32_t loc = 0;
        for (uint32_t r = t; r < c; r += T) {
            uint32_t p = alias ? (list[r].bin & 0xFFFFu) : list[r].bin;
            loc += p < 251u;
        }
        return block_sum_u32(loc, sh);
    };

    if (!bounded) {
        // FFT::compress (fft.rs:366-388): max(3, n/100) frequencies, no refinement
        uint32_t ns = nsmall(K);
        if (t == 0) {
            fw->fft_count = K;
            fw->fft_err = 0.0;
            fw->fft_size = fft_payload_size(K, ns);
            fw->fft_iters = 0;
            fw->fft_tie = (cutmask & 1u) ? TIE_FFT_TOPK : 0;
            fw->fft_valid = 1;
        }
        return;
    }

    const float Lf = (float)L;
    const double Ld = (double)L;
    auto evaluate = [&](uint32_t c) -> double {
        double acc = 0.0;
        auto epi = [&](uint32_t j, float v) {
            uint32_t ix = j < prefix ? 0u : j - prefix;  // gibbs padding replicates the edge samples
            if (ix >= N) ix = N - 1;
            const double o = d[ix];
            double out = fft_round_fast(__fdiv_rn(v, Lf), vminf, vmaxf);
            acc += mape_term(out, o);
        };
        if (gi >= 0) {
            if (sg->T4T)
                f2_inverse(*sg, ws, c, sm, epi);  // register-radix inverse (fft2.cuh)
            else
                fft_inverse(*sg, ws, c, sm, epi, d, N, prefix);
        } else {
            for (uint32_t j = t; j < N; j += T) {
                float v = 0.f;
                for (uint32_t r = 0; r < c; r++) {
                    FftEntry e = list[r];
                    if (e.bin == 0)
                        v += e.re;
                    else if (2 * e.bin == N)
                        v += (j & 1u) ? -e.re : e.re;
                    else {
                        float2 w = unit_root(e.bin * j, N, true);
                        v += 2.f * (e.re * w.x - e.im * w.y);
                    }
                }
                epi(j, v);
            }
        }
        double s = block_sum(acc, shd);
        return __ddiv_rn(s, Ld);
    };

    const int E = rust_as_i32(max_err * 1000.0);
    double cur = max_err + 1.0, prev_err = 0.0;
    uint32_t jump = 0, it = 0, c = 0, prev_c = 0xFFFFFFFFu;
    bool pruned = false, tie = false, tie_topk = false;
    while (E < rust_as_i32(cur * 1000.0)) {
        it++;
        if (!list_full && mf + jump > K) build_list(kcap);
        c = min(mf + jump, K);
        if (bound != 0xFFFFFFFFu) {
            uint32_t ns = nsmall(c);
            if (fft_payload_size(c, ns) > bound) {
                pruned = true;
                break;
            }
        }
        cur = (c == prev_c) ? prev_err : evaluate(c);
        prev_c = c;
        prev_err = cur;
        tie = tie || fft_loop_near_tie(cur, E);
        tie_topk = tie_topk || ((cutmask >> (it - 1)) & 1u);
        if (it <= 17)
            jump += hstep;
        else if (it <= 22)
            jump += tstep;
        else
            break;
    }
    uint32_t ns = nsmall(c);
    if (t == 0) {
        fw->fft_count = c;
        fw->fft_err = cur;
        fw->fft_size = fft_payload_size(c, ns);
        fw->fft_iters = (uint16_t)it;
        fw->fft_tie = (tie ? TIE_FFT_LOOP : 0) | (tie_topk ? TIE_FFT_TOPK : 0);
        fw->fft_valid = pruned ? 2 : 1;
    }
}

__global__ void __launch_bounds__(FFT_THREADS, 1) k_fft(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                               double max_err, const FftGeom *__restrict__ geoms, SlotPool pool,
                                               FftEntry *arena, float2 *spec_xd, uint32_t *spec_keys, unsigned *q) {
    extern __shared__ float2 dyn_f2[];
    __shared__ double shd[64];
    __shared__ FftGeom sg;
    __shared__ int s_item;
    FftWs ws = fft_slot(pool, blockIdx.x);
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (!fw->need_fft || fw->fft_valid == 2 || fw->fft_small) continue;  // 2: k_fft_fwd proved it cannot win
        fft_frame(samples + fw->off, fw, geoms, ws, arena + fw->fft_list_off, max_err, dyn_f2, shd, &sg, spec_xd,
                  spec_keys);
    }
}

// FFT candidate of the short frames (fft_small.cuh)
__global__ void __launch_bounds__(FS_THREADS) k_fft_small(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                                          double max_err, const FftGeom *__restrict__ geoms,
                                                          FftEntry *arena, uint32_t lmax, unsigned *q) {
    extern __shared__ double2 dyn_d2[];
    const FsSmem carved = fs_carve(reinterpret_cast<unsigned char *>(dyn_d2), lmax);
    const FsSmem *sm = &carved;
    __shared__ int s_item;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (!fw->need_fft || !fw->fft_small) continue;
        fft_small_frame(samples + fw->off, fw, geoms, arena + fw->fft_list_off, max_err, sm);
    }
}

// What a later candidate allows the FFT payload of an Auto frame to be: FFT wins only with
// size <= every passing candidate after it, and wins ties (frame/mod.rs:77,94-147).  ~0: no bound.
__device__ inline uint32_t fft_prune_bound(const FrameWork *fw, double max_err) {
    uint32_t bound = 0xFFFFFFFFu;
    if (fw->bounded && fw->comp == C_AUTO && fw->forced == 0xFF) {
        if (fw->poly_valid == 1 && fw->poly_err <= max_err) bound = min(bound, fw->poly_size);
        if (fw->need_rle) bound = min(bound, fw->rle_valid == 1 ? fw->rle_size : rle_upper_bound(fw));  // RLE's error is 0: it always passes
    }
    return bound;
}

// Probe tails of the frames k_sfold folded (sfold.cuh): small CTAs, everything after the fold in shared
// memory, only as many row pairs as the first schedule point can use (fft2.cuh: f2_probe_small).  A frame
// whose probed bins are all nonzero is settled here -- pruned (fft_valid = 2), or marked FRES_SURVIVOR so
// that k_fft_fwd transforms it without probing; the rare sparse spectrum is left to k_fft_fwd's full-row probe.
__global__ void __launch_bounds__(192, 3) k_probe(FrameWork *fr, const uint32_t *__restrict__ items, uint32_t n_items, double max_err,
                                                 const FftGeom *__restrict__ geoms, const float4 *__restrict__ fold_arena,
                                                 unsigned *q) {
    __shared__ float2 smW[2 * F2_PROBE_NP * F2_M2], smY[2 * F2_PROBE_NP * F2_M2];
    __shared__ uint32_t sh[40];
    __shared__ FftGeom sg;
    __shared__ int s_item;
    const uint32_t t = threadIdx.x;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n_items) break;
        FrameWork *fw = &fr[items[i]];
        if (!(fw->front_mode & FM_SFOLD) || !fw->need_fft || fw->f32_const || fw->geom < 0 || fw->spec_off == ~0ull) continue;
        const uint32_t bound = fft_prune_bound(fw, max_err);
        if (bound == 0xFFFFFFFFu) continue;
        if (t == 0) sg = geoms[fw->geom];
        __syncthreads();
        const uint32_t N = fw->len, mf = (3 >= N / 100) ? 3 : N / 100;
        const uint32_t want = min(mf, min(fw->fft_list_cap, (uint32_t)FFT_KCAP));  // entries of the first schedule point
        const uint32_t RB = sg.M1 / f2_fold_ra(sg.M1);
        const int np = (int)min(min((want + 485u) / 486u + 1u, (uint32_t)F2_PROBE_NP), RB);
        const uint32_t nz = block_sum_u32(f2_probe_small(fold_arena + (size_t)fw->fold_slot * SF_FOLD_SLOTS, sg, np, smW, smY), sh);
        if (nz < want) continue;  // zero bins among the probed ones: not decided here
        const uint32_t smax = sg.Bn > 65536u ? 502u : 251u;
        if (t == 0) {
            if (fft_payload_size(want, min(want, smax)) > bound) {
                fw->fft_count = want;
                fw->fft_err = max_err + 1.0;
                fw->fft_size = 0;
                fw->fft_iters = 1;
                fw->fft_tie = 0;
                fw->fft_valid = 2;  // proven unable to win
            } else {
                fw->front_res |= FRES_SURVIVOR;
            }
        }
    }
}

// forward transform (fft2.cuh) of every eligible frame, ahead of k_fft.  Auto frames whose first
// schedule point is already larger than a passing Polynomial / RLE payload end here (fft_valid = 2)
// without ever storing a spectrum; everything else leaves Xd / keys in the wave's spectrum arena.
__global__ void __launch_bounds__(F2_THREADS, 2) k_fft_fwd(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                                           double max_err, const FftGeom *__restrict__ geoms,
                                                           SlotPool pool, float2 *spec_xd, uint32_t *spec_keys,
                                                           const float4 *__restrict__ fold_arena, unsigned *q) {
    extern __shared__ float2 dyn_f2[];
    __shared__ uint32_t sh[40];
    __shared__ FftGeom sg;
    __shared__ int s_item;
    float2 *W = pool.fft_W + (size_t)blockIdx.x * MAX_FFT_LEN;
    const uint32_t t = threadIdx.x;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (!fw->need_fft || fw->f32_const || fw->geom < 0 || fw->spec_off == ~0ull) continue;
        if (fw->fft_valid == 2) continue;  // k_front's probe already proved that the candidate cannot win
        if (t == 0) sg = geoms[fw->geom];
        __syncthreads();
        const uint32_t N = fw->len;
        const bool bounded = fw->bounded != 0;
        const uint32_t prefix = (bounded && N >= 128) ? (sg.L - N) / 2 : 0u;
        float2 *Xd = spec_xd + fw->spec_off;
        uint32_t *keys = spec_keys + fw->spec_off;
        // a later candidate can only lose to FFT on size; FFT wins ties (frame/mod.rs:77,104,141)
        const uint32_t bound = fft_prune_bound(fw, max_err);
        if (bound != 0xFFFFFFFFu && (fw->front_res & FRES_SURVIVOR)) {
            // k_probe: enough nonzero bins for the first schedule point, and that point fits the bound
            f2_forward_pass1(samples + fw->off, (int)N, (int)prefix, sg, W, dyn_f2);
        } else if (bound != 0xFFFFFFFFu) {
            // The first schedule point keeps c1 = min(max_freq, #nonzero bins) entries (fft.rs:249-252),
            // at most `smax` of them with a one-byte position, and the payload only grows from there.
            // A probe of 1/8 of the spectrum (bit-identical to the full transform on those bins) usually
            // finds enough nonzero bins to prove that even this first point is larger than the bound.
            const uint32_t mf = (3 >= N / 100) ? 3 : N / 100;
            const uint32_t cap = min(fw->fft_list_cap, (uint32_t)FFT_KCAP);
            const uint32_t smax = sg.Bn > 65536u ? 502u : 251u;
            uint32_t nz;
            if (fw->front_mode & FM_SFOLD)  // k_sfold folded stage 1 while it read the frame for the stats
                nz = block_sum_u32(f2_probe_from_fold(fold_arena + (size_t)fw->fold_slot * SF_FOLD_SLOTS, sg, W, dyn_f2), sh);
            else
                nz = block_sum_u32(f2_probe(samples + fw->off, (int)N, (int)prefix, sg, W, dyn_f2), sh);
            uint32_t c1 = min(min(mf, nz), cap);
            bool pruned = fft_payload_size(c1, min(c1, smax)) > bound;
            if (!pruned) {
                // sparse spectrum (or no probe for this length): count every bin
                f2_forward_pass1(samples + fw->off, (int)N, (int)prefix, sg, W, dyn_f2);
                nz = block_sum_u32(f2_pass2<false>((int)sg.M1, sg.tw2, sg.twL1, sg.twL2, W, dyn_f2, Xd, keys), sh);
                c1 = min(min(mf, nz), cap);
                pruned = fft_payload_size(c1, min(c1, smax)) > bound;
            }
            if (pruned) {
                if (t == 0) {
                    fw->fft_count = c1;
                    fw->fft_err = max_err + 1.0;
                    fw->fft_size = 0;
                    fw->fft_iters = 1;
                    fw->fft_tie = 0;
                    fw->fft_valid = 2;
                }
                continue;
            }
        } else {
            f2_forward_pass1(samples + fw->off, (int)N, (int)prefix, sg, W, dyn_f2);
        }
        (void)f2_pass2<true>((int)sg.M1, sg.tw2, sg.twL1, sg.twL2, W, dyn_f2, Xd, keys);
        if (t == 0) fw->fwd_done = 1;
    }
}

// =========================================================================================
// noop size: Noop::optimize (noop.rs:37-43) `round() as i64`, zigzag varint
// =========================================================================================
__global__ void __launch_bounds__(BLOCK) k_noop_size(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                                     unsigned *q) {
    __shared__ double shd[64];
    __shared__ int s_item;
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (fw->comp != C_NOOP) continue;
        const double *d = samples + fw->off;
        uint32_t loc = 0;
        for (uint32_t x = threadIdx.x; x < fw->len; x += blockDim.x)
            loc += varint_len(zigzag64(rust_as_i64_safe(round(d[x]))));
        uint32_t tot = block_sum_u32(loc, (uint32_t *)shd);
        if (threadIdx.x == 0) fw->aux_size = 1 + varint_len(fw->len) + tot;
    }
}

// =========================================================================================
// selection (frame/mod.rs:71-149)
// =========================================================================================
__global__ void k_select(FrameWork *fr, uint32_t n, double max_err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameWork *fw = &fr[i];
    uint8_t winner = fw->comp, tie = 0;
    uint32_t len = 0;
    uint16_t iters = 0;
    double err = 0.0;
    const uint32_t const_size = 2 + value_bytes(fw->vmin, fw->bitdepth);
    if (fw->comp == C_AUTO) {
        if (!fw->select_only && fw->is_const) {
            winner = C_CONSTANT;
            len = const_size;
        } else if (!fw->select_only && fw->forced != 0xFF) {
            winner = fw->forced;
            if (winner == C_FFT) {
                len = fw->fft_size;
                err = fw->fft_err;
                iters = fw->fft_iters;
                tie = fw->fft_tie;
            } else if (winner == C_POLY) {
                len = fw->poly_size;
                err = fw->poly_err;
                iters = fw->poly_iters;
                tie = fw->poly_tie ? TIE_POLY_LOOP : 0;
            } else {
                len = fw->rle_size;
            }
        } else {
            // candidates in list order [FFT, Polynomial, RLE]; min_by_key keeps the first minimum;
            // candidates marked 2 were proven unable to win and are skipped
            uint32_t best = 0xFFFFFFFFu;
            if (fw->fft_valid == 1 && fw->fft_err <= max_err) {
                best = fw->fft_size;
                winner = C_FFT;
                err = fw->fft_err;
                iters = fw->fft_iters;
            }
            if (fw->poly_valid == 1 && fw->poly_err <= max_err && fw->poly_size < best) {
                best = fw->poly_size;
                winner = C_POLY;
                err = fw->poly_err;
                iters = fw->poly_iters;
            }
            if (fw->rle_valid == 1 && fw->rle_size < best) {
                best = fw->rle_size;
                winner = C_RLE;
                err = 0.0;
                iters = 0;
            }
            len = best;
            if (fw->fft_valid) {
                tie |= fw->fft_tie;
                if (fabs(fw->fft_err - max_err) < 1e-7) tie |= TIE_SELECT;
            }
            if (fw->poly_valid) {
                if (fw->poly_tie) tie |= TIE_POLY_LOOP;
                if (fabs(fw->poly_err - max_err) < 1e-12) tie |= TIE_SELECT;
            }
        }
    } else {
        switch (fw->comp) {
            case C_FFT:
                len = fw->fft_size;
                err = fw->fft_err;
                iters = fw->fft_iters;
                tie = fw->fft_tie;
                break;
            case C_POLY:
            case C_IDW:
                len = fw->poly_size;
                err = fw->poly_err;
                iters = fw->poly_iters;
                tie = fw->poly_tie ? TIE_POLY_LOOP : 0;
                break;
            case C_RLE: len = fw->rle_size; break;
            case C_CONSTANT: len = const_size; break;
            case C_NOOP: len = fw->aux_size; break;
            default: break;
        }
    }
    fw->winner = winner;
    fw->near_tie = tie;
    fw->iterations = iters;
    fw->error = err;
    fw->payload_len = fw->select_only ? 0 : len;
}

// exclusive scan of payload_len over the wave (single CTA)
__global__ void __launch_bounds__(BLOCK) k_scan(FrameWork *fr, uint32_t n, unsigned long long *total) {
    __shared__ unsigned long long s_warp[33];
    unsigned long long base = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        uint32_t i = i0 + threadIdx.x;
        unsigned long long v = i < n ? fr[i].payload_len : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long tt = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += tt;
        }
        __syncthreads();
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            unsigned long long tt = s_warp[lane], ti = tt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            s_warp[lane] = ti - tt;
            if (lane == 31) s_warp[32] = ti;
        }
        __syncthreads();
        if (i < n) fr[i].payload_off = base + s_warp[w] + inc - v;
        base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = base;
}

// =========================================================================================
// emission
// =========================================================================================
__device__ void emit_fft(const FrameWork *fw, const FftGeom *__restrict__ geoms, const FftEntry *list,
                         uint8_t *out, uint32_t *sh) {
    const uint32_t c = fw->fft_count, T = blockDim.x, t = threadIdx.x;
    const bool alias = fw->geom >= 0 && geoms[fw->geom].Bn > 65536u;
    const uint32_t hdr = 1 + varint_len(c);
    if (t == 0) {
        out[0] = 15;  // FFT_COMPRESSOR_ID (fft.rs:28)
        put_varint(out + 1, c);
    }
    uint32_t nsb = 0;  // small positions before the current tile
    for (uint32_t r0 = 0; r0 < c; r0 += T) {
        uint32_t r = r0 + t;
        FftEntry e;
        uint32_t p = 0;
        bool small = false;
        if (r < c) {
            e = list[r];
            p = alias ? (e.bin & 0xFFFFu) : e.bin;
            small = p < 251u;
        }
        uint32_t tot;
        uint32_t before = nsb + block_excl_scan_u32(small ? 1u : 0u, sh, &tot);
        if (r < c) {
            uint8_t *q = out + hdr + 11u * r - 2u * before;
            q += put_varint(q, p);
            put_bytes(q, &e.re, 4);
            put_bytes(q + 4, &e.im, 4);
        }
        nsb += tot;
        __syncthreads();
    }
    if (t == 0) {
        uint8_t *q = out + hdr + 11u * c - 2u * nsb;
        float mx = (float)fw->vmax, mn = (float)fw->vmin;
        put_bytes(q, &mx, 4);
        put_bytes(q + 4, &mn, 4);
    }
}

__device__ void emit_noop(const double *__restrict__ d, const FrameWork *fw, uint8_t *out, uint32_t *sh) {
    const uint32_t N = fw->len, T = blockDim.x, t = threadIdx.x;
    const uint32_t hdr = 1 + varint_len(N);
    if (t == 0) {
        out[0] = 250;  // NOOP_COMPRESSOR_ID (noop.rs:22)
        put_varint(out + 1, N);
    }
    uint32_t base = hdr;
    for (uint32_t x0 = 0; x0 < N; x0 += T) {
        uint32_t x = x0 + t;
        uint64_t u = 0;
        uint32_t l = 0;
        if (x < N) {
            u = zigzag64(rust_as_i64_safe(round(d[x])));
            l = varint_len(u);
        }
        uint32_t tot;
        uint32_t off = block_excl_scan_u32(l, sh, &tot);
        if (x < N) put_varint(out + base + off, u);
        base += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BLOCK) k_emit(FrameWork *fr, uint32_t n, const double *__restrict__ samples,
                                                const FftGeom *__restrict__ geoms, SlotPool pool,
                                                const FftEntry *__restrict__ arena, uint8_t *payload,
                                                const unsigned long long *total, unsigned long long cap,
                                                unsigned *overflow, unsigned *q) {
    extern __shared__ uint32_t dyn_hist[];
    __shared__ double shd[64];
    __shared__ int s_item;
    // the wave's payload (k_scan's total) must fit the device buffer the host sized before it knew
    // the total; otherwise nothing is written and the host grows the buffer and launches again
    if (*total > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1u;
        return;
    }
    uint32_t *sh = (uint32_t *)shd;
    RleWs ws = rle_slot(pool, blockIdx.x);
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        FrameWork *fw = &fr[i];
        if (fw->payload_len == 0) continue;
        const double *d = samples + fw->off;
        uint8_t *out = payload + fw->payload_off;
        switch (fw->winner) {
            case C_CONSTANT:
                if (threadIdx.x == 0) {
                    out[0] = 30;  // CONSTANT_COMPRESSOR_ID (constant.rs:26)
                    out[1] = fw->bitdepth;
                    put_value(out + 2, fw->vmin, fw->bitdepth);
                }
                break;
            case C_POLY:
            case C_IDW: poly_emit(d, fw, out, sh); break;
            case C_RLE: rle_process(d, fw, ws, out, sh, dyn_hist); break;
            case C_FFT: emit_fft(fw, geoms, arena + fw->fft_list_off, out, sh); break;
            case C_NOOP: emit_noop(d, fw, out, sh); break;
            default: break;
        }
    }
}

// =========================================================================================
// decompression (compressor/mod.rs:109-119)
// =========================================================================================
__device__ inline float load_f32(const uint8_t *p) {
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return __uint_as_float(v);
}
__device__ inline double load_f64(const uint8_t *p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return __longlong_as_double((long long)v);
}

struct DecShared {
    double dv[4];
    uint32_t u[16];
    unsigned long long m64[34];
};

// returns 0 ok, else ATSC_ERR_FORMAT-like nonzero (uniform across the CTA)
__device__ uint32_t dec_constant(const uint8_t *__restrict__ p, uint32_t len, uint32_t N, double *out,
                                 DecShared *ds) {
    if (threadIdx.x == 0) {
        uint32_t bad = 0;
        double c = 0.0;
        if (len < 3)
            bad = 1;
        else {
            uint32_t bd = p[1], l;
            if (bd == BD_U8)
                c = (double)p[2];
            else if (bd == BD_F64) {
                if (len < 10) bad = 1; else c = load_f64(p + 2);
            } else if (bd == BD_I16 || bd == BD_I32) {
                if (2 + varint_len_from_first(p[2]) > len) bad = 1;
                else {
                    int64_t v = unzigzag64(get_varint(p + 2, &l));
                    c = bd == BD_I16 ? (double)(int16_t)v : (double)(int32_t)v;
                }
            } else
                bad = 1;
        }
        ds->dv[0] = c;
        ds->u[0] = bad;
    }
    __syncthreads();
    uint32_t bad = ds->u[0];
    double c = ds->dv[0];
    __syncthreads();
    if (bad) return 5;
    // 16-byte stores over the aligned middle of the frame
    const uint32_t head = (uint32_t)(((uintptr_t)out >> 3) & 1u) & (N ? 1u : 0u), pairs = (N - head) / 2u;
    double2 *o2 = reinterpret_cast<double2 *>(out + head);
    const double2 cc = make_double2(c, c);
    for (uint32_t x = threadIdx.x; x < pairs; x += blockDim.x) o2[x] = cc;
    if (threadIdx.x == 0) {
        if (head) out[0] = c;
        if (head + 2u * pairs < N) out[N - 1] = c;
    }
    return 0;
}

__device__ uint32_t dec_noop(const uint8_t *__restrict__ p, uint32_t len, uint32_t N, double *out,
                             DecShared *ds, uint32_t *sh) {
    if (threadIdx.x == 0) {
        uint32_t bad = 0, hdr = 0;
        if (len < 2 || 1 + varint_len_from_first(p[1]) > len)
            bad = 1;
        else {
            uint32_t l;
            uint64_t cnt = get_varint(p + 1, &l);
            hdr = 1 + l;
            if (cnt != N) bad = 1;  // reference writer always stores sample_count values
        }
        ds->u[0] = bad;
        ds->u[1] = hdr;
    }
    __syncthreads();
    uint32_t bad = ds->u[0], hdr = ds->u[1];
    __syncthreads();
    if (bad) return 5;
    const uint8_t *b = p + hdr;
    uint32_t end = varscan<VarintLen>(b, len - hdr, N, sh, ds->m64, [&](uint32_t r, uint32_t off) {
        uint32_t l;
        out[r] = (double)unzigzag64(get_varint(b + off, &l));
    });
    return end == 0xFFFFFFFFu ? 5 : 0;
}

__device__ uint32_t dec_poly(const uint8_t *__restrict__ p, uint32_t len, uint32_t N, double *out, double *pts,
                             double *tang, const double *__restrict__ inv_d2, DecShared *ds, uint32_t *sh,
                             double *smem, uint32_t smem_cap) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    if (t == 0) {
        uint32_t bad = 0, hdr = 0, K = 0;
        if (len < 3 || p[0] > 1 || p[1] > 3 || 2 + varint_len_from_first(p[2]) > len)
            bad = 1;
        else {
            uint32_t l;
            uint64_t k64 = get_varint(p + 2, &l);
            hdr = 2 + l;
            if (k64 > (uint64_t)MAX_FRAME) bad = 1;
            K = (uint32_t)k64;
        }
        ds->u[0] = bad;
        ds->u[1] = hdr;
        ds->u[2] = K;
    }
    __syncthreads();
    uint32_t bad = ds->u[0], hdr = ds->u[1], K = ds->u[2];
    const uint32_t ptype = p[0], bd = p[1];
    __syncthreads();
    if (bad) return 5;
    uint32_t body;
    const uint8_t *b = p + hdr;
    if (bd == BD_U8) {
        body = K;
        if (hdr + body + 17 > len) return 5;
        for (uint32_t j = t; j < K; j += T) pts[j] = (double)b[j];
    } else if (bd == BD_F64) {
        body = 8 * K;
        if (hdr + body + 17 > len) return 5;
        for (uint32_t j = t; j < K; j += T) pts[j] = load_f64(b + 8 * (size_t)j);
    } else {
        body = varscan<VarintLen>(b, len - hdr, K, sh, ds->m64, [&](uint32_t r, uint32_t off) {
            uint32_t l;
            int64_t v = unzigzag64(get_varint(b + off, &l));
            pts[r] = bd == BD_I16 ? (double)(int16_t)v : (double)(int32_t)v;
        });
        if (body == 0xFFFFFFFFu || hdr + body + 17 > len) return 5;
    }
    __syncthreads();
    const uint8_t *tail = p + hdr + body;
    const double vmin = load_f64(tail), vmax = load_f64(tail + 8);
    const uint32_t step = tail[16];
    if (vmax == vmin) {  // polynomial.rs:396-399
        for (uint32_t x = t; x < N; x += T) out[x] = vmax;
        return 0;
    }
    if (step == 0) return 5;  // step_by(0) panics in the reference
    PolyKeys k = poly_keys(N, step);
    if (K != k.K) return 5;
    if (!ptype) {
        poly_expand(pts, k, vmin, vmax, tang, out, smem, smem_cap);
        return 0;
    }
    auto pf = [&](uint32_t j) { return pts[j]; };
    for (uint32_t x = t; x < N; x += T) out[x] = round_and_limit5(idw_eval_at(k, x, pf, inv_d2), vmin, vmax);
    return 0;
}

// rle.rs:70-110 Decode + :204-236 to_data
__device__ uint32_t dec_rle(const uint8_t *__restrict__ p, uint32_t len, uint32_t N, double *out, double *vals,
                            uint32_t *idxs, uint32_t *mark, DecShared *ds, uint32_t *sh) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    for (uint32_t x = t; x < N; x += T) mark[x] = 0;
    if (t == 0) {
        // the group structure is context dependent (value / count / indices), so one thread
        // walks it; RLE wins only when runs are few
        uint32_t bad = 0, R = 0, o = 0;
        if (len < 3 || p[1] > 3)
            bad = 1;
        else {
            uint32_t bd = p[1], l;
            o = 2;
            uint64_t U = 0;
            if (o + varint_len_from_first(p[o]) > len) bad = 1;
            else {
                U = get_varint(p + o, &l);
                o += l;
            }
            for (uint64_t g = 0; g < U && !bad; g++) {
                double v = 0.0;
                if (bd == BD_U8) {
                    if (o + 1 > len) { bad = 1; break; }
                    v = (double)p[o];
                    o += 1;
                } else if (bd == BD_F64) {
                    if (o + 8 > len) { bad = 1; break; }
                    v = load_f64(p + o);
                    o += 8;
                } else {
                    if (o >= len || o + varint_len_from_first(p[o]) > len) { bad = 1; break; }
                    int64_t iv = unzigzag64(get_varint(p + o, &l));
                    o += l;
                    v = bd == BD_I16 ? (double)(int16_t)iv : (double)(int32_t)iv;
                }
                if (o >= len || o + varint_len_from_first(p[o]) > len) { bad = 1; break; }
                uint64_t cnt = get_varint(p + o, &l);
                o += l;
                for (uint64_t k = 0; k < cnt; k++) {
                    if (o >= len || o + varint_len_from_first(p[o]) > len || R >= (uint32_t)MAX_FRAME) { bad = 1; break; }
                    uint64_t ix = get_varint(p + o, &l);
                    o += l;
                    idxs[R] = ix > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)ix;
                    vals[R] = v;
                    R++;
                }
            }
        }
        ds->u[0] = bad;
        ds->u[1] = R;
    }
    __syncthreads();
    uint32_t bad = ds->u[0], R = ds->u[1];
    __syncthreads();
    if (bad) return 5;
    for (uint32_t r = t; r < R; r += T)
        if (idxs[r] < N) atomicMax(&mark[idxs[r]], r + 1);
    __syncthreads();
    uint32_t carry = 0;
    for (uint32_t x0 = 0; x0 < N; x0 += T) {
        uint32_t x = x0 + t;
        uint32_t key = (x < N && mark[x]) ? x + 1 : 0;
        uint32_t tot;
        uint32_t pos = max(carry, block_incl_scan_max_u32(key, sh, &tot));
        if (x < N) out[x] = pos ? vals[mark[pos - 1] - 1] : 0.0;
        carry = max(carry, tot);
        __syncthreads();
    }
    return 0;
}

// fft.rs:132-144 Decode + :426-462 to_data
__device__ uint32_t dec_fft(const uint8_t *__restrict__ p, uint32_t len, uint32_t N, double *out,
                            const FftGeom *__restrict__ geoms, int gi, FftWs ws, float2 *sm, DecShared *ds,
                            uint32_t *sh, FftGeom *sg) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    if (t == 0) {
        uint32_t bad = 0, hdr = 0, c = 0;
        if (len < 10 || 1 + varint_len_from_first(p[1]) > len)
            bad = 1;
        else {
            uint32_t l;
            uint64_t c64 = get_varint(p + 1, &l);
            hdr = 1 + l;
            if (c64 > (uint64_t)FFT_DEC_KCAP) bad = 1;
            c = (uint32_t)c64;
        }
        ds->u[0] = bad;
        ds->u[1] = hdr;
        ds->u[2] = c;
    }
    __syncthreads();
    uint32_t bad = ds->u[0], hdr = ds->u[1], c = ds->u[2];
    __syncthreads();
    if (bad) return 5;
    const uint8_t *b = p + hdr;
    FftEntry *list = ws.dlist;
    uint32_t body = varscan<FftEntryLen>(b, len - hdr, c, sh, ds->m64, [&](uint32_t r, uint32_t off) {
        uint32_t l;
        FftEntry e;
        e.bin = (uint32_t)get_varint(b + off, &l) & 0xFFFFu;
        e.re = load_f32(b + off + l);
        e.im = load_f32(b + off + l + 4);
        list[r] = e;
    });
    if (body == 0xFFFFFFFFu || hdr + body + 8 > len) return 5;
    __syncthreads();
    const float vmaxf = load_f32(p + hdr + body), vminf = load_f32(p + hdr + body + 4);
    if (vmaxf == vminf) {  // fft.rs:427-430
        for (uint32_t x = t; x < N; x += T) out[x] = (double)vmaxf;
        return 0;
    }
    uint32_t L, prefix = 0;
    if (gi >= 0) {
        if (t == 0) *sg = geoms[gi];
        __syncthreads();
        L = sg->L;
        if (N >= 128) prefix = (L - N) / 2;
    } else {
        L = N;
    }
    const uint32_t half = L / 2;
    // canonicalise (pos > L/2 is the mirror of L - pos) and keep the last writer per position
    // (get_mirrored_freqs writes sequentially, fft.rs:411-420)
    for (uint32_t k = t; k <= half; k += T) ws.rank[k] = 0;
    __syncthreads();
    for (uint32_t r = t; r < c; r += T) {
        FftEntry e = list[r];
        if (e.bin >= L) {
            e.bin = 0xFFFFFFFFu;  // the reference would panic (index out of bounds)
        } else if (e.bin > half) {
            e.bin = L - e.bin;
            e.im = -e.im;
        }
        list[r] = e;
        if (e.bin != 0xFFFFFFFFu) atomicMax(&ws.rank[e.bin], r + 1);
    }
    __syncthreads();
    for (uint32_t r = t; r < c; r += T) {
        uint32_t bin = list[r].bin;
        ws.ovr[r] = bin == 0xFFFFFFFFu ? r : ws.rank[bin] - 1;  // == r when this entry is the last writer
    }
    __syncthreads();
    const float Lf = (float)L;
    auto epi = [&](uint32_t j, float v) {
        if (j >= prefix && j < prefix + N) out[j - prefix] = fft_round(__fdiv_rn(v, Lf), vminf, vmaxf);
    };
    if (gi >= 0) {
        fft_prepare_entries(*sg, ws, list, c, false, false);
        fft_inverse(*sg, ws, c, sm, epi);
    } else {
        for (uint32_t j = t; j < N; j += T) {
            float v = 0.f;
            for (uint32_t r = 0; r < c; r++) {
                FftEntry e = list[r];
                if (e.bin == 0xFFFFFFFFu || ws.ovr[r] != r) continue;
                if (e.bin == 0)
                    v += e.re;
                else if (2 * e.bin == N)
                    v += (j & 1u) ? -e.re : e.re;
                else {
                    float2 w = unit_root(e.bin * j, N, true);
                    v += 2.f * (e.re * w.x - e.im * w.y);
                }
            }
            epi(j, v);
        }
    }
    return 0;
}

__global__ void __launch_bounds__(FFT_THREADS, 2) k_decode(const DecFrame *__restrict__ fr, uint32_t n,
                                                  const uint8_t *__restrict__ payloads, double *out,
                                                  const FftGeom *__restrict__ geoms, SlotPool pool,
                                                  const double *__restrict__ inv_d2, uint32_t *status, unsigned *q) {
    extern __shared__ float2 dyn_f2[];
    __shared__ double shd[64];
    __shared__ DecShared ds;
    __shared__ FftGeom sg;
    __shared__ int s_item;
    uint32_t *sh = (uint32_t *)shd;
    FftWs fws = fft_slot(pool, blockIdx.x);
    double *pts = pool.dec_pts + (size_t)blockIdx.x * (MAX_FRAME + 8);
    double *tang = pool.poly_slope + (size_t)blockIdx.x * (MAX_FRAME + 8);
    uint32_t *mark = pool.dec_mark + (size_t)blockIdx.x * (MAX_FRAME + 8);
    uint32_t *idxs = pool.dec_idx + (size_t)blockIdx.x * (MAX_FRAME + 8);
    for (;;) {
        int i = queue_next(q, &s_item);
        if (i >= (int)n) break;
        DecFrame f = fr[i];
        const uint8_t *p = payloads + f.payload_off;
        double *o = out + f.out_off;
        uint32_t rc;
        switch (f.comp) {
            case C_CONSTANT: rc = dec_constant(p, f.payload_len, f.sample_count, o, &ds); break;
            case C_NOOP: rc = dec_noop(p, f.payload_len, f.sample_count, o, &ds, sh); break;
            case C_POLY:
            case C_IDW:
                rc = dec_poly(p, f.payload_len, f.sample_count, o, pts, tang, inv_d2, &ds, sh, reinterpret_cast<double *>(dyn_f2),
                              (uint32_t)(FFT_SMEM_BYTES / sizeof(double)));
                break;
            case C_RLE: rc = dec_rle(p, f.payload_len, f.sample_count, o, pts, idxs, mark, &ds, sh); break;
            case C_FFT: rc = dec_fft(p, f.payload_len, f.sample_count, o, geoms, f.geom, fws, dyn_f2, &ds, sh, &sg); break;
            default: rc = 4; break;
        }
        if (threadIdx.x == 0) status[i] = rc;
    }
}

__global__ void k_inv_d2(double *inv_d2, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double dd = (double)i;
        inv_d2[i] = i == 0 ? 0.0 : __ddiv_rn(1.0, __dmul_rn(dd, dd));  // 1 / d.powf(2)
    }
}

// =========================================================================================
// launchers
// =========================================================================================
static int g_sms = 0;
static int sms() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

int kernels_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, K_FFT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_front, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;

    e = cudaFuncSetAttribute(k_fft_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fft_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs_smem_bytes(FS_LMAX));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_rle, cudaFuncAttributeMaxDynamicSharedMemorySize, RLE_HIST_WORDS * 4);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, RLE_HIST_WORDS * 4);
    return (int)e;
}

static inline int grid_for(uint32_t n, int slots) { return (int)(n < (uint32_t)slots ? n : (uint32_t)slots); }

void launch_stats(const FrameWork *fr, const ChunkRef *chunks, uint32_t n_chunks, const double *samples, StatsPart *parts,
                  unsigned *q, cudaStream_t st) {
    k_stats<<<grid_for(n_chunks, 8 * sms()), 256, 0, st>>>(fr, chunks, n_chunks, samples, parts, q);
}
void launch_plan(FrameWork *fr, uint32_t n, const double *samples, const StatsPart *parts, const FftGeom *geoms,
                 P1Item *p1_list, unsigned *p1_count, cudaStream_t st) {
    k_plan<<<(n + 63) / 64, 64, 0, st>>>(fr, n, samples, parts, geoms, p1_list, p1_count);
}
void launch_poly(FrameWork *fr, uint32_t n, const double *samples, double max_err, const double *inv_d2,
                 SlotPool pool, const double *first_parts, uint32_t parts_per_item, unsigned *q, cudaStream_t st) {
    k_poly<<<grid_for(n, pool.poly_slots), 512, 0, st>>>(fr, n, samples, max_err, inv_d2, pool, first_parts, parts_per_item, q);
}
void launch_poly1(const FrameWork *fr, const ChunkRef *items, uint32_t n_items, const double *samples, double *parts,
                  unsigned *q, cudaStream_t st) {
    k_poly1<<<grid_for(n_items, 2 * sms()), 512, 0, st>>>(fr, items, n_items, samples, parts, q);
}
void launch_poly1s(const P1Item *list, const unsigned *count, uint32_t n_items, double *parts, cudaStream_t st) {
    k_poly1s<<<grid_for(n_items, 2 * sms()), P1_T, 0, st>>>(list, count, parts);
}
void launch_rle(FrameWork *fr, uint32_t n, const double *samples, double max_err, SlotPool pool, unsigned *q,
                cudaStream_t st) {
    k_rle<<<grid_for(n, pool.rle_slots), BLOCK, RLE_HIST_WORDS * 4, st>>>(fr, n, samples, max_err, pool, q);
}
void launch_fft(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                SlotPool pool, FftEntry *arena, float2 *spec_xd, uint32_t *spec_keys, unsigned *q, cudaStream_t st) {
    k_fft<<<grid_for(n, pool.fft_slots / 2), FFT_THREADS, K_FFT_SMEM_BYTES, st>>>(fr, n, samples, max_err, geoms, pool, arena,
                                                                           spec_xd, spec_keys, q);
}
void launch_fft_small(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                      FftEntry *arena, uint32_t lmax, unsigned *q, cudaStream_t st) {
    k_fft_small<<<grid_for(n, 8 * sms()), FS_THREADS, fs_smem_bytes(lmax), st>>>(fr, n, samples, max_err, geoms, arena, lmax, q);
}
void launch_fft_fwd(FrameWork *fr, uint32_t n, const double *samples, double max_err, const FftGeom *geoms,
                    SlotPool pool, float2 *spec_xd, uint32_t *spec_keys, const float4 *fold_arena, unsigned *q,
                    cudaStream_t st) {
    k_fft_fwd<<<grid_for(n, pool.fwd_slots), F2_THREADS, F2_SMEM_BYTES, st>>>(fr, n, samples, max_err, geoms, pool,
                                                                             spec_xd, spec_keys, fold_arena, q);
}
void launch_front(FrameWork *fr, const uint32_t *items, uint32_t n_items, const double *samples, double max_err,
                  const FftGeom *geoms, SlotPool pool, unsigned *q, cudaStream_t st) {
    // back-off (ns) of the producer lane's wait for a free ring slot (tunable for experiments)
    static const uint32_t nap = getenv("ATSC_FRONT_NAP") ? (uint32_t)atoi(getenv("ATSC_FRONT_NAP")) : 64u;
    k_front<<<grid_for(n_items, sms()), FR_CTA, FRONT_SMEM_BYTES, st>>>(fr, items, n_items, samples, max_err, geoms, pool, q,
                                                                        nap);
}
void launch_sfold(const FrameWork *fr, const ChunkRef *items, uint32_t n_items, const double *samples, const FftGeom *geoms,
                  float4 *fold_arena, StatsPart *parts, unsigned *q, cudaStream_t st) {
    k_sfold<<<grid_for(n_items, SF_CTAS * sms()), SF_THREADS, 0, st>>>(fr, items, n_items, samples, geoms, fold_arena, parts, q);
}
void launch_probe(FrameWork *fr, const uint32_t *items, uint32_t n_items, double max_err, const FftGeom *geoms,
                  const float4 *fold_arena, unsigned *q, cudaStream_t st) {
    k_probe<<<grid_for(n_items, 3 * sms()), 192, 0, st>>>(fr, items, n_items, max_err, geoms, fold_arena, q);
}
void launch_noop_size(FrameWork *fr, uint32_t n, const double *samples, unsigned *q, cudaStream_t st) {
    k_noop_size<<<grid_for(n, 2 * sms()), BLOCK, 0, st>>>(fr, n, samples, q);
}
void launch_select(FrameWork *fr, uint32_t n, double max_err, cudaStream_t st) {
    k_select<<<(n + 255) / 256, 256, 0, st>>>(fr, n, max_err);
}
void launch_scan(FrameWork *fr, uint32_t n, unsigned long long *total, cudaStream_t st) {
    k_scan<<<1, BLOCK, 0, st>>>(fr, n, total);
}
void launch_emit(FrameWork *fr, uint32_t n, const double *samples, const FftGeom *geoms, SlotPool pool,
                 const FftEntry *arena, uint8_t *payload, const unsigned long long *total, unsigned long long cap,
                 unsigned *overflow, unsigned *q, cudaStream_t st) {
    k_emit<<<grid_for(n, pool.rle_slots), BLOCK, RLE_HIST_WORDS * 4, st>>>(fr, n, samples, geoms, pool, arena, payload,
                                                                          total, cap, overflow, q);
}
void launch_decode(const DecFrame *fr, uint32_t n, const uint8_t *payloads, double *out, const FftGeom *geoms,
                   SlotPool pool, const double *inv_d2, uint32_t *status, unsigned *q, cudaStream_t st) {
    int slots = pool.fft_slots < pool.dec_slots ? pool.fft_slots : pool.dec_slots;
    k_decode<<<grid_for(n, slots), FFT_THREADS, FFT_SMEM_BYTES, st>>>(fr, n, payloads, out, geoms, pool, inv_d2, status, q);
}
void launch_inv_d2(double *inv_d2, uint32_t n, cudaStream_t st) { k_inv_d2<<<(n + 255) / 256, 256, 0, st>>>(inv_d2, n); }

}  // namespace atsc
