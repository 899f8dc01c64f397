// front.cuh -- k_front: the fused front end of the compress path for big frames.
//
// The reference evaluates every candidate of a frame on one in-cache slice (frame/mod.rs:112-123).
// Here a frame of >= FRONT_MIN_LEN samples is read from HBM ONCE: one CTA streams it through a
// shared-memory ring filled by bulk asynchronous copies (cp.async.bulk + mbarrier, one elected
// thread; SASS: UBLKCP / SYNCS) and, while a tile is on chip, computes
//   * DataStats + the run counts of IndexRLE      (optimizer/utils.rs:39-89, rle.rs:142-189),
//   * the MAPE of the FIRST Polynomial candidate step, which is known before the stats are
//     (step = N / max(3, N/100), polynomial.rs:209-277; utils/error.rs:104-116),
//   * stage 1 of the FFT probe: the folds over the RA contiguous chunks of the padded frame
//     (fft2.cuh: fold_acc) from which "at least c nonzero bins" is later proven without touching
//     the samples again (fft.rs:249-252; pruning rule of frame/mod.rs:94-147).
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
#pragma once
#include "common.cuh"
#include "fft2.cuh"
#include "poly.cuh"
#include "stats.cuh"

namespace atsc {

// -DFRONT_PROF: thread 0 of a few CTAs accumulates clock64() per phase and prints them at the end
#ifdef FRONT_PROF
#define FP_T(var) const long long var = clock64()
#define FP_ADD(acc, t0) (acc) += clock64() - (t0)
#else
#define FP_T(var)
#define FP_ADD(acc, t0)
#endif
struct FrontProf {
    long long wait_full = 0, pass_a = 0, barrier = 0, pass_b = 0, tail = 0, head = 0, frames = 0, tiles = 0, trips = 0, ntrips = 0, parked = 0, syncw = 0, skew = 0, issue = 0, last15 = 0, last14 = 0, lastother = 0;
};

constexpr int FR_THREADS = 512;
constexpr uint32_t FR_TILE = 4096;                     // samples per ring slot (32 KB)
constexpr uint32_t FR_SLOTS = 4;                       // two tiles pinned (scan / polynomial), two in flight
constexpr uint32_t FR_RING = FR_TILE * FR_SLOTS;       // samples; power of two
constexpr uint32_t FR_MASK = FR_RING - 1;
constexpr uint32_t FR_ROUNDS = FR_TILE / 2 / FR_THREADS;  // sample pairs per thread and tile
constexpr uint32_t FR_FOLD_SLOTS = 18 * F2_M2;         // float4 (A, B) per slot m < RB*243
constexpr uint32_t FR_NS = 4;                          // adjacent segments per thread and trip (pass B)
constexpr uint32_t FR_KMAX = 1320;                     // keys of a first step: N / 100 + 2 <= 1312
constexpr uint32_t FR_LIST = 1024;                     // parked samples per frame
constexpr uint32_t FRONT_MIN_LEN = 16384;
constexpr uint32_t FR_PRODUCER = FR_THREADS;            // first lane of the extra warp that feeds the ring
constexpr int FR_CTA = FR_THREADS + 32;                 // compute warps + the producer warp
static_assert(FR_ROUNDS * 2 * FR_THREADS == FR_TILE, "tile = whole rounds of sample pairs");

struct FrontSmem {
    double ring[FR_RING];
    float4 fold[FR_FOLD_SLOTS];
    double tang[FR_KMAX];
    uint32_t list[FR_LIST];
    double red[136];  // block reductions: 2 x 32 doubles + 4 x 32 words
    unsigned long long full[FR_SLOTS], empty[FR_SLOTS];
    unsigned long long pub_lo, pub_hi;  // ordered encodings of the range published so far
    float2 root[16];
    StatsPart part;
    uint32_t list_n;
    uint32_t varied[2];
    int item;
#ifdef FRONT_PROF
    long long arr[32];
#endif
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
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    while (!mbar_try_wait(b, parity)) {
    }
}
// global -> shared bulk copy; dst, src and bytes are multiples of 16; completion counts on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// order-preserving u64 encoding of a double (NaN never reaches it)
__device__ __forceinline__ unsigned long long ord_enc(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_dec(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e;
    return __longlong_as_double((long long)b);
}

// per-frame constants of the streaming pass
struct FrontFrame {
    const double *d;   // the frame's samples in global memory
    uint32_t N, ntiles;
    uint32_t rbase;    // ring index of sample 0
    // polynomial first step
    uint32_t step, K, Kreg, smagic;  // smagic: x / step == __umulhi(x, smagic) for x < 2^25
    // fold
    uint32_t prefix, Cc, cmagic, RA;
};

// the producer's view of a frame: what thread 0 needs to issue its tiles
struct FrontFeed {
    const double *d;
    uint32_t N, ntiles, issued;  // tiles issued so far
};

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

// ---------------------------------------------------------------------------------------
// pass A over samples [lo, hi) of tile `tile` (pointer to its first sample in the ring; lo, hi are
// tile-local and even): stats + run ends (left-neighbour form: a sample that differs from its
// predecessor ends a run at the predecessor's index) + the probe fold.  `pv_first`: the sample
// before tile-local index 0 (the previous tile's last one; sample 0 itself for the frame's first tile).
// All loads of a thread's FR_ROUNDS pairs are issued before any of them is used.
// BITS: the frame has been bit-constant so far; a warp whose samples all carry `first`'s bits skips
// the stats arithmetic (every statistic is idempotent under a repeated value).
// ---------------------------------------------------------------------------------------
template <bool FOLD, bool BITS>
__device__ __forceinline__ void front_scan(StatsAcc &a, FrontSmem *sm, const FrontFrame &f, const double *tile,
                                           uint32_t x_tile, uint32_t lo, uint32_t hi, double pv_first, double first) {
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
            pv[r] = (lo + 2u * p) ? tile[lo + 2u * p - 1u] : pv_first;
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

// pass B, one block of G * FR_NS segments: thread (g, j) takes the FR_NS ADJACENT Catmull-Rom
// segments s0 .. s0+3 at offset j (the inner keys and tangents are loaded once for the two segments
// they bound); tame arithmetic, no clamp; values outside [lo, hi] are parked.
template <bool LINEAR>
__device__ __forceinline__ void front_poly_trip(FrontSmem *sm, const FrontFrame &f, uint32_t s0, uint32_t nseg, uint32_t base,
                                                uint32_t j, double h00, double h10, double h01, double h11, double lo, double hi,
                                                double &acc) {
    const uint32_t step = f.step;
    const double *ring = sm->ring;
    double kv[FR_NS + 1], tv[FR_NS + 1], o[FR_NS], out[FR_NS], e[FR_NS];
    const double *tg = sm->tang + s0;
    if (LINEAR) {
        const double *p = ring + base, *po = p + j;
#pragma unroll
        for (uint32_t u = 0; u <= FR_NS; u++) {
            kv[u] = p[u * step];
            tv[u] = tg[u];
        }
#pragma unroll
        for (uint32_t u = 0; u < FR_NS; u++) o[u] = po[u * step];
    } else {
#pragma unroll
        for (uint32_t u = 0; u <= FR_NS; u++) {
            kv[u] = ring[(base + u * step) & FR_MASK];
            tv[u] = tg[u];
        }
#pragma unroll
        for (uint32_t u = 0; u < FR_NS; u++) o[u] = ring[(base + u * step + j) & FR_MASK];
    }
#pragma unroll
    for (uint32_t u = 0; u < FR_NS; u++) {
        const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(kv[u], h00), __dmul_rn(tv[u], h10)), __dmul_rn(kv[u + 1], h01)),
                                   __dmul_rn(tv[u + 1], h11));
        out[u] = div_1e5_int53(round_half_away(__dmul_rn(v, 100000.0)));
        e[u] = mape_term_tame(out[u], o[u]);
    }
#pragma unroll
    for (uint32_t u = 0; u < FR_NS; u++) {
        if (u < nseg) {
            if (out[u] >= lo && out[u] <= hi) {
                acc += e[u];
            } else {
                const uint32_t at = atomicAdd(&sm->list_n, 1u);
                if (at < FR_LIST) sm->list[at] = (s0 + u) * step + j;
            }
        }
    }
}

// producer thread: issues tile `feed.issued` of the frame described by `feed` into the next ring slot, as
// FR_SPLIT bulk copies that complete on the slot's barrier (one copy in flight moves only a few GB/s;
// the copy engine overlaps separate copies)
constexpr uint32_t FR_SPLIT = 2;
__device__ __forceinline__ void front_issue(FrontSmem *sm, FrontFeed &feed, uint32_t &fill) {
    const uint32_t slot = fill % FR_SLOTS, k = feed.issued;
    if (fill >= FR_SLOTS) mbar_wait(&sm->empty[slot], ((fill / FR_SLOTS) - 1u) & 1u);
    const uint32_t cnt = min(FR_TILE, feed.N - k * FR_TILE);
    mbar_expect_tx(&sm->full[slot], cnt * 8u);
    constexpr uint32_t PART = FR_TILE / FR_SPLIT;  // samples per copy (even)
    const double *src = feed.d + (size_t)k * FR_TILE;
    double *dst = sm->ring + slot * FR_TILE;
#pragma unroll
    for (uint32_t c = 0; c < FR_SPLIT; c++) {
        const uint32_t o = c * PART;
        if (o < cnt) bulk_g2s(dst + o, src + o, min(PART, cnt - o) * 8u, &sm->full[slot]);
    }
    fill++;
    feed.issued = k + 1u;
}

// One frame through the ring.  All threads of the CTA call: FR_THREADS compute threads and the
// producer warp, whose first lane (FR_PRODUCER) issues the bulk copies -- an issue costs that lane
// about a thousand cycles, which would stall a whole compute warp at every barrier.  Compute warps
// synchronise on named barrier 1 while they stream; the whole CTA meets again for the frame's tail.
// fill / use: ring slots filled / consumed since the kernel started (`fill` is the producer's).
// feed: the producer's state for THIS frame (its first tiles may already be in flight).  Having
// issued this frame's last tile the producer claims the next work item (sm->item) and goes on with
// that frame's tiles as slots come free, so the ring never drains between frames.
__device__ inline void front_frame(FrameWork *fr, const uint32_t *__restrict__ items, uint32_t n_items, int item,
                                   const double *__restrict__ samples, double max_err, const FftGeom *__restrict__ geoms,
                                   float4 *fold_arena, unsigned *q, FrontSmem *sm, uint32_t &fill, uint32_t &use,
                                   FrontFeed &feed, FrontProf &prof, uint32_t dbg) {
    const uint32_t t = threadIdx.x, lane = t & 31u;
    FP_T(t_head);
    constexpr uint32_t T = FR_THREADS;
    FrameWork *fw = &fr[items[item]];
    FrontFrame f;
    f.N = fw->len;
    f.d = samples + fw->off;
    f.ntiles = (f.N + FR_TILE - 1u) / FR_TILE;
    f.rbase = (use % FR_SLOTS) * FR_TILE;
    const uint8_t mode = fw->front_mode;
    const bool do_poly = (mode & FM_POLY) != 0;
    const bool do_fold = (mode & FM_FOLD) != 0;
    const uint32_t baseline = (3u >= f.N / 100u) ? 3u : f.N / 100u;
    f.step = max(f.N / baseline, 1u);
    f.smagic = (uint32_t)(0x100000000ull / f.step) + 1u;
    {
        const PolyKeys k = poly_keys(f.N, f.step);
        f.K = k.K;
        f.Kreg = k.Kreg;
    }
    f.prefix = f.Cc = f.cmagic = f.RA = 0;
    if (do_fold) {
        const FftGeom *g = geoms + fw->geom;
        f.RA = f2_fold_ra(g->M1);
        f.Cc = (g->M1 / f.RA) * (uint32_t)F2_M2;
        f.cmagic = (uint32_t)(0x100000000ull / f.Cc) + 1u;
        f.prefix = (g->L - f.N) / 2u;
    }
    const bool compute = t < T;
    if (t == 0) {
        sm->list_n = 0;
        sm->varied[0] = sm->varied[1] = 0;
        sm->pub_lo = ord_enc(__longlong_as_double(0x7FF0000000000000ll));
        sm->pub_hi = ord_enc(__longlong_as_double((long long)0xFFF0000000000000ull));
    }
    const double first = __ldg(f.d);
    if (do_fold) {
        if (t < 16u) {
            float2 r = make_float2(1.f, 0.f);
            switch (f.RA) {
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
        if (compute)
            for (uint32_t m = t; m < f.Cc; m += T) sm->fold[m] = m < np ? p0 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // Hermite basis of this thread's offset inside the segments (poly.cuh: poly_mape)
    const uint32_t G = T / f.step, g = t / f.step, j = t - g * f.step;
    const uint32_t BS = G * FR_NS;  // segments per block
    const double stepd = (double)f.step;
    double h00 = 0.0, h10 = 0.0, h01 = 0.0, h11 = 0.0;
    if (do_poly && g < G) {
        const double tt = __ddiv_rn((double)j, stepd);
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
    // pass-B bookkeeping, advanced block by block without divisions: the block's first segment for this
    // thread, its ring index, and the position of the key that must have landed before the block runs
    const uint32_t s_last = f.K >= 4u ? f.K - 3u : 0u;  // last Catmull-Rom segment (0: none)
    const uint32_t blk_span = BS * f.step;
    uint32_t sb = 1u;                                   // first segment of the next block
    uint32_t s0 = 1u + FR_NS * g;                       // ... and of this thread's trip in it
    uint32_t base = (s0 * f.step + f.rbase) & FR_MASK;
    uint32_t need_pos = (BS + 2u) * f.step;             // key (b+1)*BS + 2: its tangent neighbour closes the block
    uint32_t kk_next = 0;                               // next key whose arrival the tangent pass handles
    const uint32_t full0 = smem_u32(&sm->full[0]), empty0 = smem_u32(&sm->empty[0]);
    __syncthreads();  // fold / list / flags initialised
    FP_ADD(prof.head, t_head);

    // Two phases per iteration: pass A of the tile that lands now (with the tangents of the keys it
    // brings), a barrier, then pass B of every block that is complete with it.  Two tiles are pinned
    // (i-1: pass B still reads it; i), the other two ring slots are in flight.
    if (!compute) {
        // ---- producer warp: this frame's tiles, then the next frame's first ones, paced by the empty barriers
        if (t == FR_PRODUCER) {
            while (feed.issued < f.ntiles) front_issue(sm, feed, fill);
            const int nxt = (int)atomicAdd(q, 1u);
            sm->item = nxt;
            feed.issued = 0;
            feed.N = 0;
            feed.ntiles = 0;
            if (nxt < (int)n_items) {
                const FrameWork *nf = &fr[items[nxt]];
                feed.d = samples + nf->off;
                feed.N = nf->len;
                feed.ntiles = (feed.N + FR_TILE - 1u) / FR_TILE;
                // slots still in use by this frame are waited for: the producer runs as far ahead as the ring allows
                while (feed.issued < min(feed.ntiles, FR_SLOTS)) front_issue(sm, feed, fill);
            }
        }
        __syncwarp();
    } else
    for (uint32_t i = 0; i <= f.ntiles; i++) {
        const bool drain = i == f.ntiles;
        uint32_t frontier = f.N;
        if (!drain) {
            const uint32_t u = use + i, slot = u % FR_SLOTS;
            FP_T(t_w);
            {
                const uint32_t bar = full0 + slot * 8u, par = (u / FR_SLOTS) & 1u;
                uint32_t ok;
                do {
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                        "selp.u32 %0, 1, 0, p;\n"
                        "}"
                        : "=r"(ok)
                        : "r"(bar), "r"(par)
                        : "memory");
                } while (!ok);
            }
            FP_ADD(prof.wait_full, t_w);
            FP_T(t_a);
            const uint32_t xa = i * FR_TILE, xb = min(xa + FR_TILE, f.N);
            frontier = xb;
            const double *tile = sm->ring + slot * FR_TILE;
            const double pv_first = i ? sm->ring[(slot * FR_TILE - 1u) & FR_MASK] : first;
            const bool fold_now = do_fold && (seen_var || i == 0) && !(dbg & 2u);
            if (do_fold && !fold_now) skipped = true;
            // index classes of the run ends: [0, 250) | [250, 65536) | [65536, ..): 250 splits tile 0 only
            const uint32_t e0 = a.ends;
            auto scan = [&](uint32_t lo, uint32_t hi) {
                if (dbg & 1u) {
                    a.negz &= (uint32_t)__double2loint(tile[lo + 2u * t]);  // timing experiments only: touch the tile
                } else if (seen_var) {
                    if (fold_now) front_scan<true, false>(a, sm, f, tile, xa, lo, hi, pv_first, first);
                    else front_scan<false, false>(a, sm, f, tile, xa, lo, hi, pv_first, first);
                } else {
                    if (fold_now) front_scan<true, true>(a, sm, f, tile, xa, lo, hi, pv_first, first);
                    else front_scan<false, true>(a, sm, f, tile, xa, lo, hi, pv_first, first);
                }
            };
            if (xa == 0) {
                const uint32_t mid = min(250u, xb);
                scan(0, mid);
                const uint32_t e1 = a.ends;
                if (mid < xb) scan(mid, xb);
                ends250 += a.ends - e1;
            } else {
                scan(0, xb - xa);
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
            // tangents of the keys whose right-hand neighbour key landed with this tile
            if (do_poly) {
                const uint32_t kk_hi = min(__umulhi(xb - 1u, f.smagic), f.Kreg - 1u);  // last regular key in the tile
                const uint32_t kk = kk_next + (FR_THREADS - 1u - t);  // the last warps have idle lanes in pass B
                if (kk <= kk_hi && kk >= 2u && kk + 1u <= f.K) {  // tangent index kk - 1 in [1, K - 2]
                    const uint32_t pa = (kk - 2u) * f.step, pb = kk * f.step;
                    const double va = sm->ring[(pa + f.rbase) & FR_MASK], vb = sm->ring[(pb + f.rbase) & FR_MASK];
                    sm->tang[kk - 1u] = __dmul_rn(__ddiv_rn(__dsub_rn(vb, va), __dsub_rn((double)pb, (double)pa)), stepd);
                }
                kk_next = kk_hi + 1u;
                if (xb == f.N && f.K == f.Kreg + 1u && f.K >= 3u && t == 64u) {  // the appended last key N - 1
                    const uint32_t jt = f.K - 2u, pa = (jt - 1u) * f.step, pb = f.N - 1u;
                    const double va = sm->ring[(pa + f.rbase) & FR_MASK], vb = sm->ring[(pb + f.rbase) & FR_MASK];
                    sm->tang[jt] = __dmul_rn(__ddiv_rn(__dsub_rn(vb, va), __dsub_rn((double)pb, (double)pa)), stepd);
                }
            }
            FP_ADD(prof.pass_a, t_a);
        }
        FP_T(t_bar);
#ifdef FRONT_PROF
        if (lane == 0) sm->arr[t >> 5] = clock64();
#endif
        asm volatile("bar.sync 1, %0;" ::"n"(FR_THREADS) : "memory");  // compute warps only
#ifdef FRONT_PROF
        if (t == 0) {
            long long mn = sm->arr[0], mx = sm->arr[0];
            int who = 0;
            for (int k = 1; k < (int)(T >> 5); k++) {
                if (sm->arr[k] < mn) mn = sm->arr[k];
                if (sm->arr[k] > mx) { mx = sm->arr[k]; who = k; }
            }
            prof.skew += mx - mn;
            if (who == 15) prof.last15++; else if (who == 14) prof.last14++; else prof.lastother++;
        }
#endif
        FP_ADD(prof.barrier, t_bar);
        FP_T(t_b);
        if (!drain && !seen_var) seen_var = sm->varied[i & 1u] != 0u;
        // ---- pass B: every block whose last tangent exists (key (b+1)*BS + 2 landed), all at the drain
        if (do_poly && s_last) {
            if (sb <= s_last && (drain || need_pos < frontier)) {
                const double lo = ord_dec(sm->pub_lo), hi = ord_dec(sm->pub_hi);
                do {
                    if (!seen_var || (dbg & 4u)) {
                        skipped = true;
                    } else if (g < G && s0 <= s_last) {
                        const uint32_t nseg = min(FR_NS, s_last + 1u - s0);
                        FP_T(t_tr);
                        if (base + (FR_NS + 1u) * f.step < FR_RING)
                            front_poly_trip<true>(sm, f, s0, nseg, base, j, h00, h10, h01, h11, lo, hi, acc);
                        else
                            front_poly_trip<false>(sm, f, s0, nseg, base, j, h00, h10, h01, h11, lo, hi, acc);
                        FP_ADD(prof.trips, t_tr);
                        prof.ntrips += 1;
                    }
                    sb += BS;
                    s0 += BS;
                    base = (base + blk_span) & FR_MASK;
                    need_pos += blk_span;
                } while (sb <= s_last && (drain || need_pos < frontier));
            }
        }
        // release tile i-1: the unfinished blocks start within 22 * step of the frontier, inside tile i
        if (i >= 1u) {
            FP_T(t_sw);
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + ((use + i - 1u) % FR_SLOTS) * 8u) : "memory");
            FP_ADD(prof.syncw, t_sw);
        }
        FP_ADD(prof.pass_b, t_b);
    }
    FP_T(t_tail);
    prof.parked += sm->list_n;
    prof.frames += 1;
    prof.tiles += f.ntiles;
    use += f.ntiles;
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
            if (f.N > 250u && f.d[250] != f.d[249]) r1 -= 1u;
            StatsPart p;
            p.mn = mn;
            p.mx = mx;
            p.flags = fl;
            p.ends = r0;
            p.ends251 = r1;
            p.ends64k = r2;
            sm->part = p;
            finish_stats(f.d, f.N, &sm->part, 1, fw);
            plan_frame(fw);
            if (skipped && seen_var) fw->front_mode = mode & ~FM_FOLD;  // the fold misses tiles: old probe
        }
        __syncthreads();
    }
    // ======================= first Polynomial step =======================
    const double vmin = fw->vmin, vmax = fw->vmax;
    if (do_poly && fw->need_poly && fw->poly_type == 0) {
        const PolyKeys k = poly_keys(f.N, f.step);
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
        } else if (!skipped && poly_tame(vmin, vmax) && sm->list_n <= FR_LIST && f.K >= 4u && f.step < (uint32_t)POLY_MAXSTEP) {
            // parked samples: the generic arithmetic with the frame's final clamp
            auto pts = [&](uint32_t qk) { return f.d[poly_pos(k, qk)]; };
            const uint32_t nl = sm->list_n;
            for (uint32_t c = t; c < nl; c += FR_CTA) {
                const uint32_t x = sm->list[c];
                acc += mape_term(round_and_limit5_fast(poly_eval_at(k, x, pts), vmin, vmax), f.d[x]);
            }
            acc += poly_mape_ends(f.d, k, vmin, vmax);
            const double s = block_sum(acc, sm->red);
            const double cur = __ddiv_rn(s, (double)f.N);
            const double target = round_f64_dec(max_err, 3);
            const bool pass = !(target < round_f64_dec(cur, 4));  // polynomial.rs:231: the loop ends here
            uint32_t size = 0;
            if (pass) size = poly_payload_size(f.d, k, fw->bitdepth, false, reinterpret_cast<uint32_t *>(sm->red));
            if (t == 0) {
                fw->poly_step = f.step;
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
    // ======================= fold -> arena =======================
    if (do_fold && !(skipped && seen_var) && fw->need_fft) {
        __syncthreads();
        // the gibbs suffix replicates the last sample: the last chunk of the slots from (prefix + N) / 2 on
        const float lastf = (float)f.d[f.N - 1u];
        const float2 w = sm->root[f.RA - 1u];
        const uint32_t m0 = ((f.prefix + f.N) >> 1) - (f.RA - 1u) * f.Cc;
        for (uint32_t m = m0 + t; m < f.Cc; m += FR_CTA) {
            float4 ab = sm->fold[m];
            fold_acc(ab, lastf, lastf, w);
            sm->fold[m] = ab;
        }
        __syncthreads();
        float4 *dst = fold_arena + (size_t)fw->fold_idx * FR_FOLD_SLOTS;
        for (uint32_t m = t; m < f.Cc; m += FR_CTA) __stcg(dst + m, sm->fold[m]);
    }
    __syncthreads();
    FP_ADD(prof.tail, t_tail);
}

}  // namespace atsc
