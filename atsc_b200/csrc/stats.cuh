// stats.cuh -- DataStats::new (optimizer/utils.rs:39-89) + run counting for RLE
// (rle.rs:142-189), one coalesced pass over a frame by one CTA.
#pragma once
#include "common.cuh"

namespace atsc {

struct StatsSmem {
    double mn[32], mx[32];
    uint32_t flags[32], runs[32], idxb[32];
    uint32_t first_zero;
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

// Computes the frame's stats into fw (all threads must call; result written by thread 0).
// min / max follow the reference exactly: start from data[0], strict comparisons in index order
// (NaN never wins).  Equal values have equal bits except +0.0 / -0.0, so "first occurrence" only
// matters when an extreme is zero and both signs are present: that case takes a second pass.
__device__ inline void frame_stats(const double *__restrict__ d, uint32_t N, FrameWork *fw,
                                   StatsSmem *sm) {
    const double first = d[0];
    double mn = first, mx = first;
    uint32_t flags = 0;  // bit0 fractional, bit1 saw -0.0
    uint32_t runs = 0, idxb = 0;
    const uint32_t T = blockDim.x;
    const int ln = threadIdx.x & 31;
    constexpr int U = 4;  // independent coalesced loads in flight per thread
    for (uint32_t i0 = 0; i0 < N; i0 += U * T) {
        double v[U], nx[U];
        const uint32_t base = i0 + threadIdx.x;
        if (i0 + U * T <= N) {
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = d[base + u * T];
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = base + u * T < N ? d[base + u * T] : 0.0;
        }
        // the right-hand neighbour comes from a warp shuffle (only lane 31 touches memory again)
#pragma unroll
        for (int u = 0; u < U; u++) {
            nx[u] = __shfl_down_sync(0xffffffffu, v[u], 1);
            if (ln == 31 && base + u * T + 1 < N) nx[u] = d[base + u * T + 1];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = base + u * T;
            if (i >= N) continue;
            const double val = v[u];
            flags |= frac_nonzero(val) ? 1u : 0u;
            if (__double_as_longlong(val) == (long long)0x8000000000000000ull) flags |= 2u;
            if (val < mn) mn = val;
            if (val > mx) mx = val;
            // rle.rs:154: run ends where the next value differs (or at the end)
            const bool last = i + 1 >= N;
            if (last || nx[u] != val) {
                runs++;
                if (!last) idxb += 1u + ((i + 1 >= 251u) ? 2u : 0u) + ((i + 1 >= 65536u) ? 2u : 0u);  // varint_len(i + 1)
            }
        }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
        if (a < mn) mn = a;
        if (b > mx) mx = b;
        flags |= __shfl_down_sync(0xffffffffu, flags, o);
        runs += __shfl_down_sync(0xffffffffu, runs, o);
        idxb += __shfl_down_sync(0xffffffffu, idxb, o);
    }
    __syncthreads();
    if (lane == 0) {
        sm->mn[w] = mn;
        sm->mx[w] = mx;
        sm->flags[w] = flags;
        sm->runs[w] = runs;
        sm->idxb[w] = idxb;
    }
    if (threadIdx.x == 0) sm->first_zero = 0xFFFFFFFFu;
    __syncthreads();
    {
        int nw = blockDim.x >> 5;
        for (int k = 0; k < nw; k++) {
            double a = sm->mn[k], b = sm->mx[k];
            if (a < mn) mn = a;
            if (b > mx) mx = b;
            flags |= sm->flags[k];
        }
    }
    // every thread now holds the frame's min / max / flags
    double vmin = mn, vmax = mx;
    if (first != first) {
        vmin = vmax = first;  // min = max = data[0] = NaN and no comparison ever replaces it
    } else if ((flags & 2u) && (vmin == 0.0 || vmax == 0.0)) {
        // a -0.0 exists and an extreme is zero: its sign is that of the first zero in the frame
        uint32_t fz = 0xFFFFFFFFu;
        for (uint32_t x = threadIdx.x; x < N; x += T)
            if (d[x] == 0.0) {
                fz = x;
                break;
            }
        if (fz != 0xFFFFFFFFu) atomicMin(&sm->first_zero, fz);
        __syncthreads();
        const double z = d[sm->first_zero];
        if (vmin == 0.0) vmin = z;
        if (vmax == 0.0) vmax = z;
    }
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        runs = 0;
        idxb = 0;
        for (int k = 0; k < nw; k++) {
            runs += sm->runs[k];
            idxb += sm->idxb[k];
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
    __syncthreads();
}

}  // namespace atsc
