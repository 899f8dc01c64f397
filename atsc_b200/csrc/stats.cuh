// stats.cuh -- DataStats::new (optimizer/utils.rs:39-89) + run counting for RLE
// (rle.rs:142-189), one coalesced pass over a frame by one CTA.
#pragma once
#include "common.cuh"

namespace atsc {

struct StatsSmem {
    double mn[32], mx[32];
    uint32_t flags[32], runs[32], idxb[32], e64k[32];
};

// `fractional` of split_n (optimizer/utils.rs:115-160) without building the integer part.
// With e the biased exponent: e in [1023, 1074] -> some mantissa bit below the binary point is
// set; e in [959, 1022] (2^-64 <= |x| < 1) -> always; everything else (zero, denormals,
// |x| >= 2^52, inf, NaN) falls back to the bit-exact restatement (rare).
__device__ inline bool frac_nonzero(double x) {
    const uint32_t hi = (uint32_t)__double2hiint(x), lo = (uint32_t)__double2loint(x);
    const uint32_t e = (hi >> 20) & 0x7FFu;
    if (e - 1023u <= 51u) {
        // fraction bits = the low (1075 - e) bits of the 52-bit mantissa
        const uint32_t sh = e - 1023u;  // integer mantissa bits 0..51
        // mantissa << (12 + sh) as a 64-bit value != 0
        const unsigned long long m = ((unsigned long long)hi << 32 | lo) << (12u + sh);
        return m != 0ull;
    }
    if (e - 959u <= 63u) return true;
    if (e >= 1075u && e != 0x7FFu) return false;
    bool f;
    (void)split_n(x, &f);
    return f;
}

// One work item of k_stats: samples [start, start + STATS_CHUNK) of a frame.  Frames are cut into
// chunks so the pass balances over the SMs no matter how few frames a wave has.
constexpr uint32_t STATS_CHUNK = 32768;
struct ChunkRef {
    uint32_t frame, start;
};
struct StatsPart {
    double mn, mx;  // +inf / -inf when the chunk holds no comparable value
    uint32_t flags, ends, ends251, ends64k;
};

// The scan keeps the per-sample instruction count low (the pass is issue bound well before it is
// HBM bound): the right-hand neighbour comes from a plain 8-byte load at the vector load's
// address + 16 (no shuffle, no lane-31 fix-up), the fractional test is gated once per trip, the
// -0.0 test is a running unsigned minimum, and the run-end index classes need no per-sample
// counters because the caller scans each index class [0,250) [250,65535) [65535,..) on its own.
struct StatsAcc {
    double mn, mx;
    uint32_t flags;  // bit0 fractional, bit2 saw 0 < |x| < 2^-64
    uint32_t negz;   // min over samples of lo | (hi ^ 0x80000000): 0 <=> a -0.0 was seen
    uint32_t ends;   // run ends
};
__device__ __forceinline__ void stats_frac(StatsAcc &a, double val) {
    const int hi = __double2hiint(val);
    const double magic = __hiloint2double((hi & (int)0x80000000) | 0x43300000, 0);
    // fractional (split_n): r = (x + copysign(2^52, x)) - copysign(2^52, x) is x rounded to an integer
    // for |x| < 2^52 and x itself beyond (integers, inf), so "r <> x" (ordered: NaN is not
    // fractional) is exactly "x has a fractional part"; the sliver 0 < |x| < 2^-64, which split_n
    // does not call fractional, is settled by finish_stats
    const double r = __dsub_rn(__dadd_rn(val, magic), magic);
    if (r < val || r > val) a.flags |= ((uint32_t)hi & 0x7FF00000u) >= 0x3BF00000u ? 1u : 4u;
}
__device__ __forceinline__ void stats_value(StatsAcc &a, double val) {
    a.negz = min(a.negz, (uint32_t)__double2loint(val) | ((uint32_t)__double2hiint(val) ^ 0x80000000u));
    if (val < a.mn) a.mn = val;  // strict: NaN never wins (optimizer/utils.rs:57-64)
    if (val > a.mx) a.mx = val;
}
// samples dc[0 .. n): every one has a right-hand neighbour dc[i + 1] inside the frame
__device__ __forceinline__ void chunk_scan(StatsAcc &a, const double *__restrict__ dc, uint32_t n) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    uint32_t done = 0;
    if (n && (((uintptr_t)dc >> 3) & 1u)) {  // odd start: one scalar sample, then 16-byte aligned
        if (t == 0) {
            const double v = dc[0];
            if (!(a.flags & 1u)) stats_frac(a, v);
            stats_value(a, v);
            a.ends += (dc[1] != v) ? 1u : 0u;
        }
        done = 1;
    }
    const double2 *d2 = reinterpret_cast<const double2 *>(dc + done);
    const uint32_t P = (n - done) / 2;
    constexpr uint32_t U = 2;  // vector loads in flight per thread (four were no faster, at half the occupancy)
    const uint32_t Pfull = P - P % (U * T);
    uint32_t p0 = 0;
    for (; p0 < Pfull; p0 += U * T) {
        double2 v[U];
        double nx[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            const double2 *q = d2 + p0 + u * T + t;
            v[u] = __ldcs(q);                                       // streamed once
            nx[u] = __ldg(reinterpret_cast<const double *>(q + 1));  // the next pair's first sample (or dc[n])
        }
        if (!(a.flags & 1u)) {
#pragma unroll
            for (uint32_t u = 0; u < U; u++) {
                stats_frac(a, v[u].x);
                stats_frac(a, v[u].y);
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            stats_value(a, v[u].x);
            stats_value(a, v[u].y);
            a.ends += (v[u].y != v[u].x) ? 1u : 0u;
            a.ends += (nx[u] != v[u].y) ? 1u : 0u;
        }
    }
    for (uint32_t p = p0 + t; p < P; p += T) {
        const double2 *q = d2 + p;
        const double2 v = __ldcs(q);
        const double nx = __ldg(reinterpret_cast<const double *>(q + 1));
        if (!(a.flags & 1u)) {
            stats_frac(a, v.x);
            stats_frac(a, v.y);
        }
        stats_value(a, v.x);
        stats_value(a, v.y);
        a.ends += (v.y != v.x) ? 1u : 0u;
        a.ends += (nx != v.y) ? 1u : 0u;
    }
    if (t == 0 && done + 2 * P < n) {  // one leftover sample
        const double v = dc[n - 1];
        if (!(a.flags & 1u)) stats_frac(a, v);
        stats_value(a, v);
        a.ends += (dc[n] != v) ? 1u : 0u;
    }
}

// Partial stats of samples [c0, c1) of a frame of N samples (all threads call; thread 0 writes *out).
__device__ inline void chunk_stats(const double *__restrict__ d, uint32_t N, uint32_t c0, uint32_t c1, StatsPart *out,
                                    StatsSmem *sm) {
    StatsAcc a;
    a.mn = __longlong_as_double(0x7FF0000000000000ll);
    a.mx = -a.mn;
    a.flags = 0;
    a.negz = 0xFFFFFFFFu;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    // samples [c0, e) have a right-hand neighbour inside the frame; a frame's last sample is value-only
    const uint32_t e = min(c1, N - 1);
    // run-end index classes (rle.rs:160: varint_len(i + 1) is 1 / 3 / 5 bytes): i < 250, i < 65535, the rest
    uint32_t idx3[3] = {0, 0, 0};
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const uint32_t lo = max(c0, k == 0 ? 0u : k == 1 ? 250u : 65535u), hi = min(e, k == 0 ? 250u : k == 1 ? 65535u : 0xFFFFFFFFu);
        if (lo >= hi) continue;
        a.ends = 0;
        chunk_scan(a, d + lo, hi - lo);
        idx3[0] += a.ends;
        if (k >= 1) idx3[1] += a.ends;
        if (k >= 2) idx3[2] += a.ends;
    }
    if (t == 0 && c1 == N) {  // the frame's last sample
        const double v = d[N - 1];
        if (!(a.flags & 1u)) stats_frac(a, v);
        stats_value(a, v);
    }
    if (a.negz == 0u) a.flags |= 2u;
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.mn = fmin(a.mn, __shfl_down_sync(0xffffffffu, a.mn, o));
        a.mx = fmax(a.mx, __shfl_down_sync(0xffffffffu, a.mx, o));
        a.flags |= __shfl_down_sync(0xffffffffu, a.flags, o);
#pragma unroll
        for (int k = 0; k < 3; k++) idx3[k] += __shfl_down_sync(0xffffffffu, idx3[k], o);
    }
    __syncthreads();
    if (lane == 0) {
        sm->mn[w] = a.mn;
        sm->mx[w] = a.mx;
        sm->flags[w] = a.flags;
        sm->runs[w] = idx3[0];
        sm->idxb[w] = idx3[1];
        sm->e64k[w] = idx3[2];
    }
    __syncthreads();
    if (w == 0) {
        const int nw = T >> 5;
        double mn = lane < nw ? sm->mn[lane] : a.mn, mx = lane < nw ? sm->mx[lane] : a.mx;
        uint32_t fl = lane < nw ? sm->flags[lane] : 0u;
        uint32_t r0 = lane < nw ? sm->runs[lane] : 0u, r1 = lane < nw ? sm->idxb[lane] : 0u, r2 = lane < nw ? sm->e64k[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
            fl |= __shfl_down_sync(0xffffffffu, fl, o);
            r0 += __shfl_down_sync(0xffffffffu, r0, o);
            r1 += __shfl_down_sync(0xffffffffu, r1, o);
            r2 += __shfl_down_sync(0xffffffffu, r2, o);
        }
        if (lane == 0) {
            StatsPart p;
            p.mn = mn;
            p.mx = mx;
            p.flags = fl;
            p.ends = r0;
            p.ends251 = r1;
            p.ends64k = r2;
            *out = p;
        }
    }
}

// Combines a frame's chunk partials into DataStats (optimizer/utils.rs:39-89) + the run counts
// (rle.rs:142-189); one thread per frame.  min / max follow the reference exactly: start from
// data[0], strict comparisons in index order (NaN never wins).  Equal values have equal bits except
// +0.0 / -0.0, so "first occurrence" only matters when an extreme is zero and a -0.0 exists: that
// (rare) case scans for the first zero.
__device__ inline void finish_stats_with(double first, const double *__restrict__ d, uint32_t N, const StatsPart *parts,
                                         uint32_t nparts, FrameWork *fw) {
    double mn = first, mx = first;
    uint32_t flags = 0, runs = 1, idxb = 0;  // runs: + the run that ends with the last sample
    for (uint32_t k = 0; k < nparts; k++) {
        const StatsPart p = parts[k];
        if (p.mn < mn) mn = p.mn;
        if (p.mx > mx) mx = p.mx;
        flags |= p.flags;
        runs += p.ends;
        idxb += p.ends + 2u * p.ends251 + 2u * p.ends64k;
    }
    if ((flags & 5u) == 4u) {
        // no ordinary fractional value, but values below 2^-64 exist: the exact restatement decides (rare)
        for (uint32_t x = 0; x < N && !(flags & 1u); x++) {
            bool f;
            (void)split_n(d[x], &f);
            if (f) flags |= 1u;
        }
    }
    double vmin = mn, vmax = mx;
    if (first != first) {
        vmin = vmax = first;  // min = max = data[0] = NaN and no comparison ever replaces it
    } else if ((flags & 2u) && (vmin == 0.0 || vmax == 0.0)) {
        // a -0.0 exists and an extreme is zero: its sign is that of the first zero in the frame
        double z = 0.0;
        for (uint32_t x = 0; x < N; x++)
            if (d[x] == 0.0) {
                z = d[x];
                break;
            }
        if (vmin == 0.0) vmin = z;
        if (vmax == 0.0) vmax = z;
    }
    bool f;
    int64_t max_int = split_n(vmax, &f);
    int64_t min_int = split_n(vmin, &f);
    const bool frac = (flags & 1u) != 0;
    fw->vmin = vmin;
    fw->vmax = vmax;
    fw->fractional = frac ? 1 : 0;
    fw->bitdepth = frac ? BD_F64 : (uint8_t)bitdepth_of(max_int, min_int);
    fw->is_const = (vmin == vmax) ? 1 : 0;
    fw->f32_const = ((float)vmax == (float)vmin) ? 1 : 0;
    fw->n_runs = runs;
    fw->rle_idx_bytes = idxb + 1;  // + varint_len(0) for the first run
}
__device__ inline void finish_stats(const double *__restrict__ d, uint32_t N, const StatsPart *parts, uint32_t nparts,
                                    FrameWork *fw) {
    finish_stats_with(d[0], d, N, parts, nparts, fw);
}

}  // namespace atsc
