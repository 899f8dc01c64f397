// sfold.cuh -- k_sfold: DataStats + IndexRLE run counts + stage 1 of the FFT probe in ONE read.
//
// k_stats reads every sample once and k_fft_fwd's probe reads every sample of a non-constant Auto
// frame again.  Both are streaming passes with little arithmetic, so for the frames the probe
// applies to (bounded Auto frames whose padded length is 2^a * 3^7 with a >= 3: 16384 .. 131072
// samples) one pass does both.  The probe's stage 1 is DEFINED as a sequential fold over the RA
// contiguous chunks of the padded frame (fft2.cuh: fold_acc), so a thread that owns slot m walks
// the RA elements z[t * Cc + m], t = 0 .. RA-1 -- independent 16-byte loads, coalesced across the
// CTA -- keeps the fold (A, B) in four registers, runs the k_stats arithmetic on the same two
// samples, and stores the fold to the wave's fold arena (70 KB per full frame, read back from L2).
// The slots of a frame are independent, so a frame is cut into work items of SF_ITEM slots the way
// k_stats cuts it into chunks: partial stats per item, combined by k_plan.  k_fft_fwd then runs the
// rest of the probe (fold_out, stage 2, pass 2 of the 2*RB probed rows: f2_probe_from_fold) from
// the arena and applies the pruning rule without touching the samples; values are bit-identical
// to the stand-alone probe and to the full transform.
// Reference: optimizer/utils.rs:39-89, rle.rs:142-189 (stats), fft.rs:184-257 (padding, fft_trim's zero test).
#pragma once
#include "fft2.cuh"
#include "stats.cuh"

namespace atsc {

constexpr int SF_THREADS = 256;
constexpr int SF_CTAS = 4;            // per SM
constexpr int SF_U = 4;               // 16-byte loads in flight per thread

// One element of a slot's chain: the sample pair p (or a gibbs replica), stats + run ends + fold term.
//   HEAD / TAIL: the element may be a replica of the first / last sample (only chunk 0 / chunk RA-1 hold any)
//   and pair p may be the frame's last one (no right-hand neighbour); otherwise it is an interior pair.
// Run ends are counted in `ends`; `elow`: those at an index < 250 (only chunk 0 reaches down there);
// `ehi` (BIG frames, N > 65536): those of pairs >= 32768 -- the varint classes of rle.rs:160.
struct SfoldLoad {
    double2 v;
    double nx;
    int32_t p;
};
template <bool HEAD, bool TAIL>
__device__ __forceinline__ void sfold_load(SfoldLoad &l, const double2 *__restrict__ d2, int32_t p, int32_t half, double first,
                                           double last) {
    l.p = p;
    if ((HEAD && p < 0) || (TAIL && p >= half)) {
        const double c = (HEAD && p < 0) ? first : last;  // replica: idempotent for the stats, ends no run
        l.v = make_double2(c, c);
        l.nx = c;
    } else {
        const double2 *q = d2 + p;
        l.v = __ldcs(q);  // streamed once
        if (TAIL && p + 1 >= half)
            l.nx = 0.0;  // the last pair has no right-hand neighbour (not used below)
        else
            l.nx = __ldg(reinterpret_cast<const double *>(q + 1));
    }
}
template <bool HEAD, bool TAIL, bool BIG>
__device__ __forceinline__ void sfold_use(StatsAcc &a, uint32_t &elow, uint32_t &ehi, const SfoldLoad &l, int32_t half) {
    stats_value(a, l.v.x);
    stats_value(a, l.v.y);
    uint32_t e = (l.v.y != l.v.x) ? 1u : 0u;
    if (HEAD || TAIL) {
        const bool real = !(HEAD && l.p < 0) && !(TAIL && l.p >= half);  // a replica may be NaN
        e = real ? e : 0u;
        e += (real && !(TAIL && l.p + 1 >= half) && l.nx != l.v.y) ? 1u : 0u;
    } else {
        e += (l.nx != l.v.y) ? 1u : 0u;
    }
    a.ends += e;
    if (HEAD && l.p < 125) elow += e;
    if (BIG && l.p >= 32768) ehi += e;
}

// Slots [s0, s1) of one frame: samples d[0 .. N), N even, d 16-byte aligned.  Slot m of the padded frame's
// complex elements (np = prefix / 2 replicas of sample 0 in front, replicas of the last sample behind,
// fft.rs:184-204) is the chain z[t * Cc + m], t = 0 .. RA-1.
template <int RA, bool BIG>
__device__ __forceinline__ void sfold_stream(StatsAcc &a, uint32_t &elow, uint32_t &ehi, float4 *__restrict__ fold,
                                             const double *__restrict__ d, uint32_t N, uint32_t np, uint32_t Cc, uint32_t s0,
                                             uint32_t s1, double first, double last) {
    constexpr int U = RA < SF_U ? RA : SF_U;
    constexpr int NG = RA / U;
    const int32_t half = (int32_t)(N >> 1);
    const double2 *d2 = reinterpret_cast<const double2 *>(d);
    for (uint32_t m = s0 + threadIdx.x; m < s1; m += SF_THREADS) {
        float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
        const int32_t p0 = (int32_t)m - (int32_t)np;  // pair of the chunk-0 element
#pragma unroll
        for (int g = 0; g < NG; g++) {
            SfoldLoad l[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int t = g * U + u;
                const int32_t p = p0 + (int32_t)((uint32_t)t * Cc);
                if (t == 0) sfold_load<true, false>(l[u], d2, p, half, first, last);
                else if (t == RA - 1) sfold_load<false, true>(l[u], d2, p, half, first, last);
                else sfold_load<false, false>(l[u], d2, p, half, first, last);
            }
            if (!(a.flags & 1u)) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    stats_frac(a, l[u].v.x);
                    stats_frac(a, l[u].v.y);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int t = g * U + u;
                if (t == 0) sfold_use<true, false, BIG>(a, elow, ehi, l[u], half);
                else if (t == RA - 1) sfold_use<false, true, BIG>(a, elow, ehi, l[u], half);
                else sfold_use<false, false, BIG>(a, elow, ehi, l[u], half);
                fold_acc(ab, (float)l[u].v.x, (float)l[u].v.y, root_c<RA, false>(t));
            }
        }
        fold[m] = ab;
    }
}

// One work item: slots [s0, s0 + SF_ITEM) of a frame; all threads call, thread 0 writes the item's
// partial stats (combined by k_plan: finish_stats).
__device__ inline void sfold_item(const FrameWork *fw, uint32_t s0, const double *__restrict__ samples,
                                  const FftGeom *__restrict__ geoms, float4 *__restrict__ fold_arena, StatsPart *__restrict__ parts,
                                  StatsSmem *sm) {
    const uint32_t t = threadIdx.x, N = fw->len;
    const double *d = samples + fw->off;
    const double first = __ldg(d), last = __ldg(d + N - 1);
    const FftGeom *g = geoms + fw->geom;
    const uint32_t M1 = g->M1, RA = f2_fold_ra(M1), Cc = (M1 / RA) * (uint32_t)F2_M2, np = (g->L - N) >> 2;
    const uint32_t s1 = min(s0 + SF_ITEM, Cc);
    float4 *fold = fold_arena + (size_t)fw->fold_slot * SF_FOLD_SLOTS;
    StatsAcc a;
    a.mn = __longlong_as_double(0x7FF0000000000000ll);
    a.mx = -a.mn;
    a.flags = 0;
    a.negz = 0xFFFFFFFFu;
    a.ends = 0;
    uint32_t elow = 0, ehi = 0;  // run ends at an index < 250; of pairs >= 32768
    switch (RA) {
        case 16:
            if (N > 65536u) sfold_stream<16, true>(a, elow, ehi, fold, d, N, np, Cc, s0, s1, first, last);
            else sfold_stream<16, false>(a, elow, ehi, fold, d, N, np, Cc, s0, s1, first, last);
            break;
        case 8: sfold_stream<8, false>(a, elow, ehi, fold, d, N, np, Cc, s0, s1, first, last); break;
        default: sfold_stream<4, false>(a, elow, ehi, fold, d, N, np, Cc, s0, s1, first, last); break;
    }
    if (a.negz == 0u) a.flags |= 2u;
    uint32_t r0 = a.ends, r1 = elow, r2 = ehi;
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.mn = fmin(a.mn, __shfl_down_sync(0xffffffffu, a.mn, o));
        a.mx = fmax(a.mx, __shfl_down_sync(0xffffffffu, a.mx, o));
        a.flags |= __shfl_down_sync(0xffffffffu, a.flags, o);
        r0 += __shfl_down_sync(0xffffffffu, r0, o);
        r1 += __shfl_down_sync(0xffffffffu, r1, o);
        r2 += __shfl_down_sync(0xffffffffu, r2, o);
    }
    __syncthreads();
    if (lane == 0) {
        sm->mn[w] = a.mn;
        sm->mx[w] = a.mx;
        sm->flags[w] = a.flags;
        sm->runs[w] = r0;
        sm->idxb[w] = r1;
        sm->e64k[w] = r2;
    }
    __syncthreads();
    if (t == 0) {
        double mn = sm->mn[0], mx = sm->mx[0];
        uint32_t fl = sm->flags[0];
        r0 = sm->runs[0];
        r1 = sm->idxb[0];
        r2 = sm->e64k[0];
        for (int k = 1; k < SF_THREADS / 32; k++) {
            mn = fmin(mn, sm->mn[k]);
            mx = fmax(mx, sm->mx[k]);
            fl |= sm->flags[k];
            r0 += sm->runs[k];
            r1 += sm->idxb[k];
            r2 += sm->e64k[k];
        }
        StatsPart p;
        p.mn = mn;
        p.mx = mx;
        p.flags = fl;
        p.ends = r0;
        p.ends251 = r0 - r1;  // ends at an index >= 250
        // ends at an index >= 65535: the pairs from 32768 on and, once per frame, the second sample of pair 32767
        p.ends64k = r2 + ((s0 == 0 && N > 65536u && d[65536] != d[65535]) ? 1u : 0u);
        parts[fw->chunk0 + s0 / SF_ITEM] = p;
    }
}

}  // namespace atsc
