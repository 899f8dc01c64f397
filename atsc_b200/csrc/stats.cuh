// stats.cuh -- DataStats::new (optimizer/utils.rs:39-89) + run counting for RLE
// (rle.rs:142-189), one coalesced pass over a frame by one CTA.
#pragma once
#include "common.cuh"

namespace atsc {

struct MinMaxIdx {
    double v;
    uint32_t i;  // 0xFFFFFFFF = empty
};

// "first occurrence of the smallest": strict compare on value, then lower index
__device__ inline MinMaxIdx mm_min(MinMaxIdx a, MinMaxIdx b) {
    if (a.i == 0xFFFFFFFFu) return b;
    if (b.i == 0xFFFFFFFFu) return a;
    if (b.v < a.v) return b;
    if (a.v < b.v) return a;
    return a.i <= b.i ? a : b;
}
__device__ inline MinMaxIdx mm_max(MinMaxIdx a, MinMaxIdx b) {
    if (a.i == 0xFFFFFFFFu) return b;
    if (b.i == 0xFFFFFFFFu) return a;
    if (b.v > a.v) return b;
    if (a.v > b.v) return a;
    return a.i <= b.i ? a : b;
}
__device__ inline MinMaxIdx mm_shfl_down(MinMaxIdx a, int o) {
    MinMaxIdx r;
    r.v = __shfl_down_sync(0xffffffffu, a.v, o);
    r.i = __shfl_down_sync(0xffffffffu, a.i, o);
    return r;
}

struct StatsSmem {
    MinMaxIdx mn[32], mx[32];
    uint32_t frac[32], runs[32], idxb[32];
};

// Computes the frame's stats into fw (all threads must call; result written by thread 0).
__device__ inline void frame_stats(const double *__restrict__ d, uint32_t N, FrameWork *fw,
                                   StatsSmem *sm) {
    MinMaxIdx mn = {0.0, 0xFFFFFFFFu}, mx = {0.0, 0xFFFFFFFFu};
    uint32_t frac = 0, runs = 0, idxb = 0;
    // 4 independent coalesced loads in flight per thread; the right-hand neighbour comes from a
    // warp shuffle (only lane 31 touches memory again)
    const uint32_t T = blockDim.x;
    const int ln = threadIdx.x & 31;
    for (uint32_t i0 = 0; i0 < N; i0 += 4 * T) {
        double v[4], nx[4];
        uint32_t ix[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            ix[u] = i0 + u * T + threadIdx.x;
            v[u] = ix[u] < N ? d[ix[u]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            nx[u] = __shfl_down_sync(0xffffffffu, v[u], 1);
            if (ln == 31 && ix[u] + 1 < N) nx[u] = d[ix[u] + 1];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (ix[u] >= N) continue;
            const uint32_t i = ix[u];
            const double val = v[u];
            bool f;
            (void)split_n(val, &f);
            frac |= f ? 1u : 0u;
            if (val == val) {  // NaN never wins a strict comparison (optimizer/utils.rs:57-64)
                if (mn.i == 0xFFFFFFFFu || val < mn.v) {
                    mn.v = val;
                    mn.i = i;
                }
                if (mx.i == 0xFFFFFFFFu || val > mx.v) {
                    mx.v = val;
                    mx.i = i;
                }
            }
            // rle.rs:154: run ends where the next value differs (or at the end)
            bool end = (i + 1 >= N) || (nx[u] != val);
            if (end) {
                runs++;
                if (i + 1 < N) idxb += varint_len((uint64_t)i + 1);
            }
        }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = mm_min(mn, mm_shfl_down(mn, o));
        mx = mm_max(mx, mm_shfl_down(mx, o));
        frac |= __shfl_down_sync(0xffffffffu, frac, o);
        runs += __shfl_down_sync(0xffffffffu, runs, o);
        idxb += __shfl_down_sync(0xffffffffu, idxb, o);
    }
    __syncthreads();
    if (lane == 0) {
        sm->mn[w] = mn;
        sm->mx[w] = mx;
        sm->frac[w] = frac;
        sm->runs[w] = runs;
        sm->idxb[w] = idxb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; k++) {
            mn = mm_min(mn, sm->mn[k]);
            mx = mm_max(mx, sm->mx[k]);
            frac |= sm->frac[k];
            runs += sm->runs[k];
            idxb += sm->idxb[k];
        }
        double first = d[0];
        double vmin, vmax;
        if (first != first) {
            // min = max = data[0] = NaN and no comparison ever replaces it
            vmin = vmax = first;
        } else {
            vmin = mn.v;
            vmax = mx.v;
        }
        bool f;
        int64_t max_int = split_n(vmax, &f);
        int64_t min_int = split_n(vmin, &f);
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
