// front.cuh -- k_front: the fused front end of the compress path for big frames.
//
// The reference evaluates every candidate of a frame on one in-cache slice (frame/mod.rs:112-123).
// Here a frame of >= FRONT_MIN_LEN samples is read from HBM ONCE: one CTA streams it through a
// shared-memory ring filled by bulk asynchronous copies (cp.async.bulk + mbarrier, issued by a
// producer warp; SASS: UBLKCP / SYNCS) and, while a tile is on chip, computes
//   * DataStats + the run counts of IndexRLE      (optimizer/utils.rs:39-89, rle.rs:142-189),
//   * the MAPE of the FIRST Polynomial candidate step, which is known before the stats are
//     (step = N / max(3, N/100) = 100 for every frame this kernel takes; polynomial.rs:209-277,
//     utils/error.rs:104-116),
//   * the FFT probe: stage 1 (the folds over the RA contiguous chunks of the padded frame, fft2.cuh:
//     fold_acc) while streaming, the rest at the end of the frame, which proves "at least c nonzero
//     bins" without touching the samples again (fft.rs:249-252; pruning rule of frame/mod.rs:94-147).
// A frame whose Polynomial candidate passes at its first step and whose FFT candidate is pruned --
// every big frame of a monitoring fleet -- never sees k_poly / the sample-reading probe.
//
// What cannot be known while streaming is the frame's [min, max] (the clamp of
// round_and_limit_f64, utils/mod.rs:66-74) and whether the frame is "tame" (poly.cuh).  The sum is
// therefore accumulated with the tame arithmetic and WITHOUT the clamp for every value inside the
// range seen so far (an inner bound of the final range: the clamp cannot act there); the few
// values outside it are parked in a list and added, properly clamped, at the end of the frame.  A
// frame that turns out not to be tame, overflows the list, or started with constant tiles whose
// work was skipped falls back to k_poly / the old probe: same results, one more read.
//
// Ring: FR_SLOTS slots of [halo | tile].  A tile is 40 segments of the first step; the Polynomial
// pass of a tile lags the stats pass by at most three segments (the tangent of a key needs the next
// key), and those samples are copied into the NEXT slot's halo before the tile is released.  So only
// the tile being worked on is pinned, three are in flight, and every address is linear.
#pragma once
#include "common.cuh"
#include "fft2.cuh"
#include "poly.cuh"
#include "rle.cuh"
#include "stats.cuh"

namespace atsc {

constexpr int FR_THREADS = 512;                        // compute threads
constexpr int FR_CTA = FR_THREADS + 32;                // + the producer warp
constexpr uint32_t FR_PRODUCER = FR_THREADS;           // its first lane issues the bulk copies
constexpr uint32_t FR_STEP = 100;                      // first Polynomial step of every frame k_front takes
constexpr uint32_t FR_TILE = 40 * FR_STEP;             // samples per tile (32,000 B)
constexpr uint32_t FR_HALO = 512;                      // samples kept from the previous tile (>= 3 segments)
constexpr uint32_t FR_SLOT = FR_HALO + FR_TILE;        // samples per ring slot
constexpr uint32_t FR_SLOTS = 4;
constexpr uint32_t FR_ROUNDS = (FR_TILE / 2 + FR_THREADS - 1) / FR_THREADS;  // sample pairs per thread and tile
constexpr uint32_t FR_G = FR_THREADS / FR_STEP;        // groups of FR_STEP threads (pass B)
constexpr uint32_t FR_NS = 4;                          // adjacent segments per thread and trip (pass B)
constexpr uint32_t FR_TW = 512;                        // rolling window of keys / tangents
constexpr uint32_t FR_FOLD_SLOTS = 18 * F2_M2;         // float4 (A, B) per slot m < RB*243
constexpr uint32_t FR_LIST = 1024;                     // parked samples per frame
constexpr uint32_t FRONT_MIN_LEN = 16384;
static_assert(FR_HALO >= 3 * FR_STEP + 4 && FR_HALO % 2 == 0, "halo covers the polynomial lag");
static_assert((FR_SLOT * 8) % 16 == 0 && (FR_HALO * 8) % 16 == 0, "bulk copies are 16-byte aligned");

// what the producer warp prepares for a frame while the previous one streams
struct FrontDesc {
    const double *d;   // the frame's samples in global memory
    double first;      // sample 0
    uint32_t idx;      // work item (>= n_items: none left)
    uint32_t frame;    // index into the FrameWork table
    uint32_t N, ntiles, mode;
    uint32_t prefix, Cc, cmagic, RA;  // fold geometry (FM_FOLD)
};

struct FrontSmem {
    double ring[FR_SLOTS * FR_SLOT];
    float4 fold[FR_FOLD_SLOTS];
    double tang[FR_TW + 8];  // entries 0..7 mirrored behind the window: five consecutive reads never wrap
    double keyv[FR_TW];
    uint32_t list[FR_LIST];
    double red[136];  // block reductions: 2 x 32 doubles + 4 x 32 words
    unsigned long long full[FR_SLOTS], empty[FR_SLOTS];
    unsigned long long pub_lo, pub_hi;  // ordered encodings of the range published so far
    float2 root[16];
    FrontDesc desc[2];
    StatsPart part;
    double s249, s250, lastv;      // samples captured while streaming
    double fin_vmin, fin_vmax;     // the frame's final range, for the whole CTA
    uint32_t fin_flags;            // bit0 need_poly (Catmull-Rom), bit1 need_fft, bit2 work was skipped, bit3 ... and missed, bits 8.. bitdepth
    uint32_t list_n;
    uint32_t varied[2];
};
constexpr int FRONT_SMEM_BYTES = (int)sizeof(FrontSmem);
static_assert(sizeof(FrontSmem) <= 227 * 1024, "k_front shared memory");

// ---------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX; the ring is CTA-local)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; dst, src and bytes are multiples of 16; completion counts on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// the compute warps' barrier (the producer warp never joins it)
__device__ __forceinline__ void front_bar() { asm volatile("bar.sync 1, %0;" ::"n"(FR_THREADS) : "memory"); }

// order-preserving u64 encoding of a double (NaN never reaches it)
__device__ __forceinline__ unsigned long long ord_enc(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_dec(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
    return __longlong_as_double((long long)b);
}

// which candidates run for a frame whose stats are final (frame/mod.rs:71-149, compressor/mod.rs:63-107);
// shared by k_plan and k_front
__device__ inline void plan_frame(FrameWork *fw) {
    uint8_t np = 0, nr = 0, nf = 0, pt = 0;
    switch (fw->comp) {
        case C_AUTO:
            if (fw->select_only || (!fw->is_const && fw->forced == 0xFF))
                np = nr = nf = 1;
            else if (!fw->is_const) {
                // frame/mod.rs:106-111: the sampled pass chose; compress everything with it
                nf = fw->forced == C_FFT;
                np = fw->forced == C_POLY;
                nr = fw->forced == C_RLE;
            }
            break;
        case C_FFT: nf = 1; break;
        case C_POLY: np = 1; break;
        case C_IDW:
            np = 1;
            pt = 1;
            break;
        case C_RLE: nr = 1; break;
        default: break;
    }
    fw->need_poly = np;
    fw->need_rle = nr;
    fw->need_fft = nf;
    fw->poly_type = pt;
    fw->poly_valid = fw->rle_valid = fw->fft_valid = 0;
    fw->poly_size = fw->rle_size = fw->fft_size = 0;
    fw->poly_err = fw->fft_err = 0.0;
    fw->poly_tie = fw->fft_tie = 0;
    fw->poly_iters = fw->fft_iters = 0;
    fw->fft_count = 0;
    fw->aux_size = 0;
    fw->fwd_done = 0;
    fw->front_res = 0;
}

// per-frame constants of the compute threads
struct FrontFrame {
    uint32_t N, K, Kreg;
    uint32_t prefix, Cc, cmagic;  // fold
};

// ---------------------------------------------------------------------------------------
// pass A over samples [lo, hi) of a tile (tile-local, even): stats + run ends (left-neighbour form:
// a sample that differs from its predecessor ends a run at the predecessor's index) + the probe
// fold.  `pv0`: the sample before tile-local index 0 (only the thread of pair 0 uses it).
// All loads of a thread's FR_ROUNDS pairs are issued before any of them is used.
// BITS: the frame has been bit-constant so far; a warp whose samples all carry `first`'s bits skips
// the stats arithmetic (every statistic is idempotent under a repeated value).
// ---------------------------------------------------------------------------------------
template <bool FOLD, bool BITS>
__device__ __forceinline__ void front_scan(StatsAcc &a, FrontSmem *sm, const FrontFrame &f, const double *tile, uint32_t x_tile,
                                           uint32_t lo, uint32_t hi, double pv0, double first) {
    const uint32_t t = threadIdx.x;
    const uint32_t P = (hi - lo) >> 1;
    const double2 *tp = reinterpret_cast<const double2 *>(tile + lo);
    double2 v[FR_ROUNDS];
    double pv[FR_ROUNDS];
#pragma unroll
    for (uint32_t r = 0; r < FR_ROUNDS; r++) {
        const uint32_t p = t + r * FR_THREADS;
        if (p < P) {
            v[r] = tp[p];
            pv[r] = (lo + 2u * p) ? tile[lo + 2u * p - 1u] : pv0;
        } else {
            v[r] = make_double2(first, first);
            pv[r] = first;
        }
    }
    bool work = true;
    if (BITS) {
        const uint32_t flo = (uint32_t)__double2loint(first), fhi = (uint32_t)__double2hiint(first);
        uint32_t diff = 0;
#pragma unroll
        for (uint32_t r = 0; r < FR_ROUNDS; r++) {
            diff |= ((uint32_t)__double2loint(v[r].x) ^ flo) | ((uint32_t)__double2hiint(v[r].x) ^ fhi);
            diff |= ((uint32_t)__double2loint(v[r].y) ^ flo) | ((uint32_t)__double2hiint(v[r].y) ^ fhi);
            diff |= ((uint32_t)__double2loint(pv[r]) ^ flo) | ((uint32_t)__double2hiint(pv[r]) ^ fhi);
        }
        work = __any_sync(0xffffffffu, diff != 0u);
    }
    if (work) {
        if (!(a.flags & 1u)) {
#pragma unroll
            for (uint32_t r = 0; r < FR_ROUNDS; r++) {  // padding lanes repeat `first`: idempotent
                stats_frac(a, v[r].x);
                stats_frac(a, v[r].y);
            }
        }
#pragma unroll
        for (uint32_t r = 0; r < FR_ROUNDS; r++) {
            stats_value(a, v[r].x);
            stats_value(a, v[r].y);
            if (t + r * FR_THREADS < P) {
                a.ends += (pv[r] != v[r].x) ? 1u : 0u;
                a.ends += (v[r].x != v[r].y) ? 1u : 0u;
            }
        }
    }
    if (FOLD) {
        const uint32_t n0 = (f.prefix + x_tile + lo) >> 1;  // complex element of the padded frame of pair 0
#pragma unroll
        for (uint32_t r = 0; r < FR_ROUNDS; r++) {
            const uint32_t p = t + r * FR_THREADS;
            if (p < P) {
                const uint32_t n = n0 + p;
                const uint32_t tc = __umulhi(n, f.cmagic);  // n / Cc: its chunk ...
                const uint32_t m = n - tc * f.Cc;           // ... and slot
                float4 ab = sm->fold[m];
                fold_acc(ab, (float)v[r].x, (float)v[r].y, sm->root[tc]);
                sm->fold[m] = ab;
            }
        }
    }
}

// values of the first step that the clamp cannot touch go into the sum; the others are parked
__device__ __forceinline__ void front_take(FrontSmem *sm, double out, double e, double lo, double hi, uint32_t x, double &acc) {
    if (out >= lo && out <= hi) {
        acc += e;
    } else {
        const uint32_t at = atomicAdd(&sm->list_n, 1u);
        if (at < FR_LIST) sm->list[at] = x;
    }
}

// pass B, one trip: thread (g, j) takes the FR_NS ADJACENT Catmull-Rom segments s0 .. s0+3 at offset j
// (the inner keys and tangents are loaded once for the two segments they bound); `p` points at key s0
// in the ring (halo or tile: linear).  Tame arithmetic, no clamp.
__device__ __forceinline__ void front_poly_trip(FrontSmem *sm, const double *p, uint32_t s0, uint32_t nseg, uint32_t j, double h00,
                                                double h10, double h01, double h11, double lo, double hi, double &acc) {
    double kv[FR_NS + 1], tv[FR_NS + 1], o[FR_NS], out[FR_NS], e[FR_NS];
    const double *tg = sm->tang + (s0 & (FR_TW - 1u)), *po = p + j;
#pragma unroll
    for (uint32_t u = 0; u <= FR_NS; u++) {
        kv[u] = p[u * FR_STEP];
        tv[u] = tg[u];
    }
#pragma unroll
    for (uint32_t u = 0; u < FR_NS; u++) o[u] = po[u * FR_STEP];
#pragma unroll
    for (uint32_t u = 0; u < FR_NS; u++) {
        const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(kv[u], h00), __dmul_rn(tv[u], h10)), __dmul_rn(kv[u + 1], h01)),
                                   __dmul_rn(tv[u + 1], h11));
        out[u] = div_1e5_int53(round_half_away(__dmul_rn(v, 100000.0)));
        e[u] = mape_term_tame(out[u], o[u]);
    }
#pragma unroll
    for (uint32_t u = 0; u < FR_NS; u++)
        if (u < nseg) front_take(sm, out[u], e[u], lo, hi, (s0 + u) * FR_STEP + j, acc);
}

// producer lane: issues tile `issued` of the frame described by `D` into the next ring slot.  It sits in
// the highest-numbered warp, which the issue arbiter prefers: its wait backs off (`nap` ns) so that a hot
// spin does not take issue slots from the compute warps of its sub-partition.
__device__ __forceinline__ void front_issue(FrontSmem *sm, const FrontDesc &D, uint32_t &issued, uint32_t &fill, uint32_t full0,
                                            uint32_t empty0, uint32_t nap) {
    const uint32_t slot = fill % FR_SLOTS, k = issued;
    if (fill >= FR_SLOTS)
        while (!mbar_try_wait(empty0 + slot * 8u, ((fill / FR_SLOTS) - 1u) & 1u))
            if (nap) __nanosleep(nap);
    const uint32_t cnt = min(FR_TILE, D.N - k * FR_TILE);
    mbar_expect_tx(full0 + slot * 8u, cnt * 8u);
    bulk_g2s(sm->ring + slot * FR_SLOT + FR_HALO, D.d + (size_t)k * FR_TILE, cnt * 8u, full0 + slot * 8u);
    fill++;
    issued = k + 1u;
}

// producer lane: claims the next work item and describes its frame in *D
__device__ inline void front_claim(FrontDesc *D, const FrameWork *fr, const uint32_t *__restrict__ items, uint32_t n_items,
                                   const double *__restrict__ samples, const FftGeom *__restrict__ geoms, unsigned *q) {
    const uint32_t idx = atomicAdd(q, 1u);
    D->idx = idx;
    D->N = D->ntiles = 0;
    if (idx >= n_items) return;
    const uint32_t fi = items[idx];
    const FrameWork *fw = &fr[fi];
    D->frame = fi;
    D->d = samples + fw->off;
    D->N = fw->len;
    D->ntiles = (fw->len + FR_TILE - 1u) / FR_TILE;
    D->mode = fw->front_mode;
    D->prefix = D->Cc = D->cmagic = D->RA = 0;
    if (fw->front_mode & FM_FOLD) {
        const FftGeom *g = geoms + fw->geom;
        D->RA = f2_fold_ra(g->M1);
        D->Cc = (g->M1 / D->RA) * (uint32_t)F2_M2;
        D->cmagic = (uint32_t)(0x100000000ull / D->Cc) + 1u;
        D->prefix = (g->L - fw->len) / 2u;
    }
    D->first = __ldg(D->d);
}

// One frame through the ring.  All threads of the CTA call: FR_THREADS compute threads and the
// producer warp, whose first lane (FR_PRODUCER) issues the bulk copies -- an issue costs that lane
// about a thousand cycles, which would stall a whole compute warp at every barrier.  Compute warps
// synchronise on named barrier 1 while they stream; the whole CTA meets again for the frame's tail.
// fill / use: ring slots filled / consumed since the kernel started (`fill`, `issued` are the
// producer's).  Having issued this frame's last tile the producer claims the next work item
// (sm->desc[(fc + 1) & 1]) and goes on with that frame's tiles as slots come free, so the ring
// never drains between frames.
__device__ inline void front_frame(FrameWork *fr, const uint32_t *__restrict__ items, uint32_t n_items, uint32_t fc,
                                   const double *__restrict__ samples, double max_err, const FftGeom *__restrict__ geoms,
                                   float2 *W, unsigned *q, FrontSmem *sm, uint32_t &fill, uint32_t &use, uint32_t &issued,
                                   uint32_t nap) {
    const uint32_t t = threadIdx.x, lane = t & 31u;
    constexpr uint32_t T = FR_THREADS;
    const bool compute = t < T;
    const FrontDesc &D = sm->desc[fc & 1u];
    FrameWork *fw = &fr[D.frame];
    const double *d = D.d;
    FrontFrame f;
    f.N = D.N;
    const uint32_t ntiles = D.ntiles;
    const uint8_t mode = (uint8_t)D.mode;
    const bool do_poly = (mode & FM_POLY) != 0;
    const bool do_fold = (mode & FM_FOLD) != 0;
    {
        const PolyKeys k = poly_keys(f.N, FR_STEP);
        f.K = k.K;
        f.Kreg = k.Kreg;
    }
    f.prefix = D.prefix;
    f.Cc = D.Cc;
    f.cmagic = D.cmagic;
    const uint32_t RA = D.RA;
    const double first = D.first;
    const uint32_t full0 = smem_u32(&sm->full[0]), empty0 = smem_u32(&sm->empty[0]);
    if (t == 0) {
        sm->list_n = 0;
        sm->varied[0] = sm->varied[1] = 0;
        sm->pub_lo = ord_enc(__longlong_as_double(0x7FF0000000000000ll));
        sm->pub_hi = ord_enc(__longlong_as_double((long long)0xFFF0000000000000ull));
    }
    if (do_fold && compute) {
        if (t < 16u) {
            float2 r = make_float2(1.f, 0.f);
            switch (RA) {
                case 16: r = root_c<16, false>((int)t); break;
                case 8: r = root_c<8, false>((int)(t & 7u)); break;
                default: r = root_c<4, false>((int)(t & 3u)); break;
            }
            sm->root[t] = r;
        }
        // the gibbs prefix replicates sample 0 (fft.rs:184-204): chunk 0 of the slots below prefix / 2
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f);
        fold_acc(p0, (float)first, (float)first, make_float2(1.f, 0.f));
        const uint32_t np = f.prefix >> 1;
        for (uint32_t m = t; m < f.Cc; m += T) sm->fold[m] = m < np ? p0 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // Hermite basis of this thread's offset inside the segments (poly.cuh: poly_mape)
    const uint32_t g = t / FR_STEP, j = t - g * FR_STEP;
    constexpr double stepd = (double)FR_STEP;
    double tt = 0.0, h00 = 0.0, h10 = 0.0, h01 = 0.0, h11 = 0.0;
    if (do_poly && g < FR_G) {
        tt = __ddiv_rn((double)j, stepd);
        const double two_t = __dmul_rn(tt, 2.0), three_t = __dmul_rn(tt, 3.0);
        const double t2 = __dmul_rn(tt, tt), t3 = __dmul_rn(t2, tt);
        const double two_t3 = __dmul_rn(t2, two_t), two_t2 = __dmul_rn(tt, two_t), three_t2 = __dmul_rn(tt, three_t);
        h00 = __dadd_rn(__dsub_rn(two_t3, three_t2), 1.0);
        h10 = __dadd_rn(__dsub_rn(t3, two_t2), tt);
        h01 = __dsub_rn(three_t2, two_t3);
        h11 = __dsub_rn(t3, t2);
    }
    StatsAcc a;
    a.mn = __longlong_as_double(0x7FF0000000000000ll);
    a.mx = -a.mn;
    a.flags = 0;
    a.negz = 0xFFFFFFFFu;
    // left-neighbour form compares sample 0 with itself: a NaN there would count a run end that does not exist
    a.ends = (t == 0 && !(first == first)) ? 0xFFFFFFFFu : 0u;
    stats_frac(a, first);  // sample 0 once in every thread: a warp of bit-identical samples may then skip
    stats_value(a, first);
    uint32_t ends250 = 0, ends64k = 0;  // run ends at index >= 250 / >= 65535 (rle.rs:160 varint classes)
    double pub_mn = __longlong_as_double(0x7FF0000000000000ll), pub_mx = -pub_mn;  // what this warp last published
    double acc = 0.0;
    bool seen_var = !(first == first);  // a NaN first sample: min = max = NaN, never "constant"
    bool skipped = false;
    const uint32_t s_last = f.K >= 4u ? f.K - 3u : 0u;  // last Catmull-Rom segment (0: none)
    uint32_t s_next = 1u;                               // first Catmull-Rom segment not yet evaluated
    double prev_last = first;                           // thread 0: the sample before the current tile
    __syncthreads();  // fold / list / flags initialised

    if (!compute) {
        // ---- producer warp: this frame's tiles, then the next frame's first ones, paced by the empty barriers
        if (t == FR_PRODUCER) {
            while (issued < ntiles) front_issue(sm, D, issued, fill, full0, empty0, nap);
            FrontDesc *nd = &sm->desc[(fc + 1u) & 1u];
            front_claim(nd, fr, items, n_items, samples, geoms, q);
            issued = 0;
            while (issued < min(nd->ntiles, FR_SLOTS)) front_issue(sm, *nd, issued, fill, full0, empty0, nap);
        }
        __syncwarp();
    } else {
        for (uint32_t i = 0; i < ntiles; i++) {
            const uint32_t u = use + i, slot = u % FR_SLOTS;
            const uint32_t xa = i * FR_TILE, xb = min(xa + FR_TILE, f.N);
            const double *tile = sm->ring + slot * FR_SLOT + FR_HALO;
            mbar_wait(full0 + slot * 8u, (u / FR_SLOTS) & 1u);
            // ---- pass A: stats, run ends by index class [0, 250) | [250, 65536) | [65536, ..), fold
            const bool fold_now = do_fold && (seen_var || i == 0);
            if (do_fold && !fold_now) skipped = true;
            auto scan = [&](uint32_t lo, uint32_t hi) {  // frame indices, even
                if (lo >= hi) return;
                if (seen_var) {
                    if (fold_now) front_scan<true, false>(a, sm, f, tile, xa, lo - xa, hi - xa, prev_last, first);
                    else front_scan<false, false>(a, sm, f, tile, xa, lo - xa, hi - xa, prev_last, first);
                } else {
                    if (fold_now) front_scan<true, true>(a, sm, f, tile, xa, lo - xa, hi - xa, prev_last, first);
                    else front_scan<false, true>(a, sm, f, tile, xa, lo - xa, hi - xa, prev_last, first);
                }
            };
            if (xa < 250u || (xa < 65536u && xb > 65536u)) {  // a class boundary inside the tile (tile 0; one more tile)
                scan(xa, min(xb, 250u));
                uint32_t e0 = a.ends;
                scan(max(xa, 250u), min(xb, 65536u));
                ends250 += a.ends - e0;
                e0 = a.ends;
                scan(max(xa, 65536u), xb);
                ends250 += a.ends - e0;
                ends64k += a.ends - e0;
                if (xa == 0 && t == 1u && xb > 250u) {
                    sm->s249 = tile[249];
                    sm->s250 = tile[250];
                }
            } else {
                const uint32_t e0 = a.ends;
                scan(xa, xb);
                ends250 += a.ends - e0;
                if (xa >= 65536u) ends64k += a.ends - e0;
            }
            // publish the range seen so far (inner bound of the frame's range) and whether anything varied
            if (__any_sync(0xffffffffu, a.mn < pub_mn || a.mx > pub_mx)) {
                double mn = a.mn, mx = a.mx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                pub_mn = mn;
                pub_mx = mx;
                if (lane == 0) {
                    atomicMin(&sm->pub_lo, ord_enc(mn));
                    atomicMax(&sm->pub_hi, ord_enc(mx));
                }
                if (!seen_var && (mn < first || mx > first)) sm->varied[i & 1u] = 1u;
            }
            // keys of this tile and the tangents they complete: tangent kk-1 = (key kk - key kk-2) / 200 * 100
            if (do_poly) {
                const uint32_t kk_lo = (xa + FR_STEP - 1u) / FR_STEP, kk_hi = min((xb - 1u) / FR_STEP, f.Kreg - 1u);
                const uint32_t kk = kk_lo + (T - 1u - t);  // the last warps have idle lanes in pass B
                if (kk <= kk_hi) {
                    const double vk = tile[kk * FR_STEP - xa];
                    sm->keyv[kk & (FR_TW - 1u)] = vk;
                    if (kk >= 2u && kk + 1u <= f.K) {  // tangent index kk - 1 in [1, K - 2]
                        const uint32_t pa = (kk - 2u) * FR_STEP;
                        const double va = pa >= xa ? tile[pa - xa] : sm->keyv[(kk - 2u) & (FR_TW - 1u)];
                        const double tg = __dmul_rn(__ddiv_rn(__dsub_rn(vk, va), 2.0 * stepd), stepd);
                        const uint32_t ti = (kk - 1u) & (FR_TW - 1u);
                        sm->tang[ti] = tg;
                        if (ti < 8u) sm->tang[ti + FR_TW] = tg;
                    }
                }
                if (xb == f.N && f.K == f.Kreg + 1u && t == 64u) {  // the appended last key N - 1 closes tangent K - 2
                    const uint32_t jt = f.K - 2u, pa = (jt - 1u) * FR_STEP, pb = f.N - 1u;
                    const double va = pa >= xa ? tile[pa - xa] : sm->keyv[(jt - 1u) & (FR_TW - 1u)], vb = tile[pb - xa];
                    const double tg = __dmul_rn(__ddiv_rn(__dsub_rn(vb, va), __dsub_rn((double)pb, (double)pa)), stepd);
                    const uint32_t ti = jt & (FR_TW - 1u);
                    sm->tang[ti] = tg;
                    if (ti < 8u) sm->tang[ti + FR_TW] = tg;
                }
            }
            front_bar();
            if (!seen_var) seen_var = sm->varied[i & 1u] != 0u;
            // ---- pass B: the Catmull-Rom segments whose right-hand tangent exists (key s + 2 landed)
            if (do_poly && s_last) {
                const uint32_t s_max = xb == f.N ? s_last : min(s_last, (xb - 1u) / FR_STEP - 2u);
                const double lo = ord_dec(sm->pub_lo), hi = ord_dec(sm->pub_hi);
                if (!seen_var) {
                    skipped = true;
                } else {
                    if (g < FR_G) {
                        const double *p0 = tile + (int)(s_next * FR_STEP) - (int)xa;  // key s_next: in the halo for the first segments
                        for (uint32_t s0 = s_next + FR_NS * g; s0 <= s_max; s0 += FR_NS * FR_G)
                            front_poly_trip(sm, p0 + (s0 - s_next) * FR_STEP, s0, min(FR_NS, s_max + 1u - s0), j, h00, h10, h01, h11,
                                            lo, hi, acc);
                    }
                    // the Linear ends (polynomial.rs:349): segment 0 with the first tile, segment K-2 and the
                    // last sample with the last one -- same parking rule
                    if (i == 0 && g == 0) {
                        const double v = __dadd_rn(__dmul_rn(tile[0], __dsub_rn(1.0, tt)), __dmul_rn(tile[FR_STEP], tt));
                        const double out = div_1e5_int53(round_half_away(__dmul_rn(v, 100000.0)));
                        front_take(sm, out, mape_term_tame(out, tile[j]), lo, hi, j, acc);
                    }
                    if (xb == f.N) {
                        const uint32_t start_last = (f.K - 2u) * FR_STEP, x = start_last + (T - 1u - t);
                        if (x < f.N) {
                            double v = tile[f.N - 1u - xa];
                            if (x != f.N - 1u) {
                                const double at = (double)start_last, bt = (double)(f.N - 1u);
                                const double nt = __ddiv_rn(__dsub_rn((double)x, at), __dsub_rn(bt, at));
                                const double ka = tile[(int)start_last - (int)xa];
                                v = __dadd_rn(__dmul_rn(ka, __dsub_rn(1.0, nt)), __dmul_rn(v, nt));
                            }
                            const double out = div_1e5_int53(round_half_away(__dmul_rn(v, 100000.0)));
                            front_take(sm, out, mape_term_tame(out, tile[(int)x - (int)xa]), lo, hi, x, acc);
                        }
                    }
                }
                s_next = s_max + 1u;
            }
            // ---- the next tile's halo, the carry of thread 0, release
            if (xb < f.N) {
                if (t < FR_HALO / 2u) {
                    const double2 hv = *reinterpret_cast<const double2 *>(tile + FR_TILE - FR_HALO + 2u * t);
                    *reinterpret_cast<double2 *>(sm->ring + ((u + 1u) % FR_SLOTS) * FR_SLOT + 2u * t) = hv;
                }
            } else if (t == 0) {
                sm->lastv = tile[f.N - 1u - xa];
            }
            if (t == 0) prev_last = tile[xb - xa - 1u];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + slot * 8u);
        }
    }
    use += ntiles;

    // ======================= end of the frame: stats =======================
    {
        if (a.negz == 0u) a.flags |= 2u;
        double mn = a.mn, mx = a.mx;
        uint32_t fl = a.flags, r0 = a.ends, r1 = ends250, r2 = ends64k;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
            fl |= __shfl_down_sync(0xffffffffu, fl, o);
            r0 += __shfl_down_sync(0xffffffffu, r0, o);
            r1 += __shfl_down_sync(0xffffffffu, r1, o);
            r2 += __shfl_down_sync(0xffffffffu, r2, o);
        }
        uint32_t *ru = reinterpret_cast<uint32_t *>(sm->red + 64);
        const uint32_t w = t >> 5;
        if (lane == 0) {
            sm->red[w] = mn;
            sm->red[32 + w] = mx;
            ru[w] = fl;
            ru[32 + w] = r0;
            ru[64 + w] = r1;
            ru[96 + w] = r2;
        }
        __syncthreads();
        if (t == 32u) {
            mn = sm->red[0];
            mx = sm->red[32];
            fl = ru[0];
            r0 = ru[32];
            r1 = ru[64];
            r2 = ru[96];
            for (uint32_t k = 1; k < (T >> 5); k++) {  // compute warps (the producer warp saw sample 0 only)
                mn = fmin(mn, sm->red[k]);
                mx = fmax(mx, sm->red[32 + k]);
                fl |= ru[k];
                r0 += ru[32 + k];
                r1 += ru[64 + k];
                r2 += ru[96 + k];
            }
            // the end at index 249 was counted with the [250, ..) class (left-neighbour form)
            if (f.N > 250u && sm->s250 != sm->s249) r1 -= 1u;
            StatsPart p;
            p.mn = mn;
            p.mx = mx;
            p.flags = fl;
            p.ends = r0;
            p.ends251 = r1;
            p.ends64k = r2;
            sm->part = p;
            finish_stats_with(first, d, f.N, &sm->part, 1, fw);
            plan_frame(fw);
            if (skipped && seen_var) fw->front_mode = mode & ~FM_FOLD;  // the fold misses tiles: old probe
            sm->fin_vmin = fw->vmin;
            sm->fin_vmax = fw->vmax;
            // decisions for the whole CTA (the producer warp did not stream: its `skipped` / `seen_var` say nothing)
            sm->fin_flags = ((fw->need_poly && fw->poly_type == 0) ? 1u : 0u) | (fw->need_fft ? 2u : 0u) | (skipped ? 4u : 0u) |
                            ((skipped && seen_var) ? 8u : 0u) | ((uint32_t)fw->bitdepth << 8);
        }
        __syncthreads();
    }
    // ======================= first Polynomial step =======================
    const double vmin = sm->fin_vmin, vmax = sm->fin_vmax;
    const uint32_t fin = sm->fin_flags;
    if (do_poly && (fin & 1u)) {
        const PolyKeys k = poly_keys(f.N, FR_STEP);
        if (vmax == vmin) {
            // polynomial.rs:210,280 "Same max and min": no points
            if (t == 0) {
                fw->poly_step = 1;
                fw->poly_npts = 0;
                fw->poly_size = 1 + 1 + 1 + 16 + 1;
                fw->poly_iters = 0;
                fw->poly_tie = 0;
                fw->poly_err = 0.0;
                fw->poly_valid = 1;
            }
        } else if (!(fin & 4u) && poly_tame(vmin, vmax) && sm->list_n <= FR_LIST && f.K >= 4u) {
            // parked samples: the generic arithmetic with the frame's final clamp
            auto pts = [&](uint32_t qk) { return d[poly_pos(k, qk)]; };
            const uint32_t nl = sm->list_n;
            for (uint32_t c = t; c < nl; c += FR_CTA) {
                const uint32_t x = sm->list[c];
                acc += mape_term(round_and_limit5_fast(poly_eval_at(k, x, pts), vmin, vmax), d[x]);
            }
            const double s = block_sum(acc, sm->red);
            const double cur = __ddiv_rn(s, (double)f.N);
            const double target = round_f64_dec(max_err, 3);
            const bool pass = !(target < round_f64_dec(cur, 4));  // polynomial.rs:231: the loop ends here
            uint32_t size = 0;
            if (pass) size = poly_payload_size(d, k, (int)((fin >> 8) & 0xFFu), false, reinterpret_cast<uint32_t *>(sm->red));
            if (t == 0) {
                fw->poly_step = FR_STEP;
                fw->poly_err = cur;
                if (pass) {
                    fw->poly_npts = k.K;
                    fw->poly_size = size;
                    fw->poly_iters = 1;
                    fw->poly_tie = poly_loop_near_tie(cur, target) ? 1 : 0;
                    fw->poly_valid = 1;
                } else {
                    fw->front_res |= FRES_POLY1;  // k_poly goes on from the second step
                }
            }
        }
    }
    // ======================= FFT probe from the fold =======================
    // (the rule of k_fft_fwd: the first schedule point keeps c1 = min(max_freq, #nonzero bins) entries,
    // fft.rs:249-252, and the payload only grows from there; FFT wins only with size <= the passing
    // candidates after it, frame/mod.rs:94-147)
    if (do_fold && !(fin & 8u) && (fin & 2u)) {
        // the gibbs suffix replicates the last sample: the last chunk of the slots from (prefix + N) / 2 on
        const float lastf = (float)sm->lastv;
        const float2 w = sm->root[RA - 1u];
        const uint32_t m0 = ((f.prefix + f.N) >> 1) - (RA - 1u) * f.Cc;
        for (uint32_t m = m0 + t; m < f.Cc; m += FR_CTA) {
            float4 ab = sm->fold[m];
            fold_acc(ab, lastf, lastf, w);
            sm->fold[m] = ab;
        }
        __syncthreads();
        const FftGeom &G = geoms[fw->geom];
        // stage 2 reads the fold and writes the probe intermediate to W (L2); pass 2 then reuses the fold's memory
        const uint32_t nzl = f2_probe_from_fold(sm->fold, G, W, reinterpret_cast<float2 *>(sm->fold));
        const uint32_t nz = block_sum_u32(nzl, reinterpret_cast<uint32_t *>(sm->red));
        if (t == 0) {
            uint32_t bound = 0xFFFFFFFFu;
            if (fw->poly_valid == 1 && fw->poly_err <= max_err) bound = min(bound, fw->poly_size);
            if (fw->need_rle) bound = min(bound, rle_upper_bound(fw));  // RLE's error is 0: it always passes
            const uint32_t mf = (3u >= f.N / 100u) ? 3u : f.N / 100u;
            const uint32_t cap = min(fw->fft_list_cap, (uint32_t)FFT_KCAP);
            const uint32_t smax = G.Bn > 65536u ? 502u : 251u;
            const uint32_t c1 = min(min(mf, nz), cap);
            if (bound != 0xFFFFFFFFu && fft_payload_size(c1, min(c1, smax)) > bound) {
                fw->fft_count = c1;
                fw->fft_err = max_err + 1.0;
                fw->fft_size = 0;
                fw->fft_iters = 1;
                fw->fft_tie = 0;
                fw->fft_valid = 2;  // proven unable to win
            } else {
                fw->front_mode = mode & ~FM_FOLD;  // sparse spectrum: k_fft_fwd counts every bin
            }
        }
    }
    __syncthreads();
}

}  // namespace atsc
