// rle.cuh -- index RLE (rle.rs:142-189 IndexRLE::new, :40-67 Encode), device side.
// The reference's BTreeMap<u64 bits, Vec<usize>> is a group-by over run starts ordered by
// f64::to_bits; here: ordered compaction of runs, stable LSD radix sort by the 64-bit key
// (8-bit digits, uniform bytes skipped), boundary detection, size / emission by prefix sums --
// all as warp-chunked ordered passes (one block-wide scan per pass, not one per 1024 elements).
#pragma once
#include "common.cuh"

namespace atsc {

struct RleWs {
    uint64_t *k0, *k1;  // run keys (ping-pong), capacity MAX_FRAME each
    uint32_t *i0, *i1;  // run start indices (ping-pong)
    uint32_t *bnd;      // scratch: run ends, later group boundary positions (capacity MAX_FRAME + 1)
};

constexpr int RLE_HIST_WORDS = 16 * BLOCK;  // dynamic smem words needed by rle_process

// lower bound of the RLE payload size without grouping (used to prune the sort in Auto mode)
__device__ inline uint32_t rle_lower_bound(const FrameWork *fw) {
    uint32_t minvb = fw->bitdepth == BD_F64 ? 8u : 1u;
    uint32_t groups = fw->is_const ? 1u : 2u;
    return 2u + 1u + groups * (minvb + 1u) + fw->rle_idx_bytes;
}

// upper bound of the RLE payload size without grouping: every run its own group (a merged group
// of c runs costs one value + varint(c) instead of c values + c count bytes, never more).  The
// FFT candidate, which runs BEFORE the RLE sort, is pruned against this figure.
__device__ inline uint32_t rle_upper_bound(const FrameWork *fw) {
    const uint32_t maxvb = fw->bitdepth == BD_F64 ? 8u : fw->bitdepth == BD_I32 ? 5u : fw->bitdepth == BD_I16 ? 3u : 1u;
    const unsigned long long ub = 2ull + 5ull + (unsigned long long)fw->n_runs * (maxvb + 1u) + fw->rle_idx_bytes;
    return ub > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)ub;
}

// ---------------------------------------------------------------------------------------
// Warp-chunked ordered passes.  Every pass over an array of n elements gives warp w the
// contiguous chunk [w*chunk, (w+1)*chunk) and walks it 32 elements at a time, so element order
// == (warp, step, lane) order and prefix quantities need one block-wide scan of 32 warp totals
// instead of one per 1024 elements.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rle_chunk(uint32_t n, uint32_t nwarps) {
    uint32_t c = (n + nwarps - 1) / nwarps;
    return (c + 31u) & ~31u;
}
// exclusive scan of one value per warp (lane 0 holds it); every thread of warp w gets warp w's
// prefix, *total the sum.  scratch: >= 33 words.
__device__ inline uint32_t rle_warp_offsets(uint32_t v, uint32_t *scratch, uint32_t *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    uint32_t mine = lane < nw ? scratch[lane] : 0u, inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t excl = __shfl_sync(0xffffffffu, inc - mine, w);
    *total = tot;
    __syncthreads();
    return excl;
}

// All threads of the CTA call.  Returns the exact payload size; if out != nullptr also
// writes the payload.  sh: 128-word static scratch; hist: RLE_HIST_WORDS words.
__device__ inline uint32_t rle_process(const double *__restrict__ d, FrameWork *fw, RleWs ws,
                                       uint8_t *out, uint32_t *sh, uint32_t *hist) {
    const uint32_t N = fw->len;
    const int bitdepth = fw->bitdepth;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t lane = t & 31u, w = t >> 5, W = T >> 5;
    const uint32_t lt = (1u << lane) - 1u;

    // ---- A. ordered compaction of run ends (rle.rs:154: a run ends where the next value differs)
    const uint32_t cA = rle_chunk(N, W), a0 = min(N, w * cA), a1 = min(N, (w + 1) * cA);
    auto run_end = [&](uint32_t i, double &v) -> bool {
        v = i < a1 ? d[i] : 0.0;
        double nx = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 31 && i + 1 < N) nx = d[i + 1];
        return i < a1 && (i + 1 >= N || nx != v);
    };
    uint32_t cnt = 0;
    for (uint32_t i0 = a0; i0 < a1; i0 += 32) {
        double v;
        cnt += __popc(__ballot_sync(0xffffffffu, run_end(i0 + lane, v)));
    }
    uint32_t R;
    uint32_t pos = rle_warp_offsets(cnt, sh, &R);
    for (uint32_t i0 = a0; i0 < a1; i0 += 32) {
        double v;
        const bool end = run_end(i0 + lane, v);
        const uint32_t m = __ballot_sync(0xffffffffu, end);
        if (end) {
            const uint32_t r = pos + __popc(m & lt);
            ws.k0[r] = (uint64_t)__double_as_longlong(v);  // bits of the run's value
            ws.bnd[r] = i0 + lane;
        }
        pos += __popc(m);
    }
    __syncthreads();
    // ---- B. run starts
    for (uint32_t r = t; r < R; r += T) ws.i0[r] = r == 0 ? 0u : ws.bnd[r - 1] + 1u;
    // ---- C. which key bytes vary
    uint64_t key0 = ws.k0[0];
    uint32_t dlo = 0, dhi = 0;
    for (uint32_t r = t; r < R; r += T) {
        uint64_t x = ws.k0[r] ^ key0;
        dlo |= (uint32_t)x;
        dhi |= (uint32_t)(x >> 32);
    }
    __syncthreads();
    if (t == 0) {
        sh[100] = 0;
        sh[101] = 0;
    }
    __syncthreads();
    if (dlo) atomicOr(&sh[100], dlo);
    if (dhi) atomicOr(&sh[101], dhi);
    __syncthreads();
    const uint64_t diff = ((uint64_t)sh[101] << 32) | sh[100];
    __syncthreads();

    // ---- D. stable LSD radix sort on the varying bytes.  hist[w*256 + digit]: first per-warp
    // digit counts, then (after one scan in digit-major, warp-minor order) each warp's running
    // destination per digit; a warp's chunk is contiguous, so walking it in order keeps the sort stable.
    uint64_t *ks = ws.k0, *kd = ws.k1;
    uint32_t *is = ws.i0, *id = ws.i1;
    const uint32_t cR = rle_chunk(R, W), r0w = min(R, w * cR), r1w = min(R, (w + 1) * cR);
    uint32_t *hw = hist + w * 256u;
    for (int p = 0; p < 8; p++) {
        if (((diff >> (8 * p)) & 255ull) == 0) continue;
        for (uint32_t q = t; q < W * 256u; q += T) hist[q] = 0;
        __syncthreads();
        for (uint32_t rb = r0w; rb < r1w; rb += 32) {
            const uint32_t r = rb + lane;
            const bool in = r < r1w;
            const uint32_t dg = in ? (uint32_t)(ks[r] >> (8 * p)) & 255u : 256u + lane;  // idle lanes: unique
            const uint32_t m = __match_any_sync(0xffffffffu, dg);
            if (in && (m & lt) == 0) hw[dg] += __popc(m);  // lowest lane of each digit group
            __syncwarp();
        }
        __syncthreads();
        {
            // linear index L = digit * W + warp; thread t owns L = 8t .. 8t+7 (W == 32: one digit, 8 warps)
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t L = t * 8 + e, dg = L / W, ww = L - dg * W;
                loc[e] = (L < W * 256u) ? hist[ww * 256u + dg] : 0u;
                sum += loc[e];
            }
            uint32_t tot;
            uint32_t ex = block_excl_scan_u32(sum, sh, &tot);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t L = t * 8 + e, dg = L / W, ww = L - dg * W;
                if (L < W * 256u) hist[ww * 256u + dg] = ex;
                ex += loc[e];
            }
        }
        __syncthreads();
        for (uint32_t rb = r0w; rb < r1w; rb += 32) {
            const uint32_t r = rb + lane;
            const bool in = r < r1w;
            const uint64_t key = in ? ks[r] : 0ull;
            const uint32_t dg = in ? (uint32_t)(key >> (8 * p)) & 255u : 256u + lane;
            const uint32_t m = __match_any_sync(0xffffffffu, dg);
            uint32_t base = in ? hw[dg] : 0u;
            __syncwarp();
            if (in && (m & lt) == 0) hw[dg] = base + __popc(m);
            __syncwarp();
            if (in) {
                const uint32_t dst = base + __popc(m & lt);
                kd[dst] = key;
                id[dst] = is[r];
            }
        }
        __syncthreads();
        uint64_t *tk = ks;
        ks = kd;
        kd = tk;
        uint32_t *ti = is;
        is = id;
        id = ti;
    }

    // ---- E. groups: positions where the sorted key changes -> bnd[g]; group index of every run -> gidx[r]
    uint32_t *gidx = id;  // the idle ping-pong index buffer
    auto grp_start = [&](uint32_t r) -> bool { return r < r1w && (r == 0 || ks[r] != ks[r - 1]); };
    cnt = 0;
    for (uint32_t rb = r0w; rb < r1w; rb += 32) cnt += __popc(__ballot_sync(0xffffffffu, grp_start(rb + lane)));
    uint32_t U;
    uint32_t gpos = rle_warp_offsets(cnt, sh, &U);
    for (uint32_t rb = r0w; rb < r1w; rb += 32) {
        const uint32_t r = rb + lane;
        const bool b = grp_start(r);
        const uint32_t m = __ballot_sync(0xffffffffu, b);
        const uint32_t g = gpos + __popc(m & lt) + (b ? 1u : 0u) - 1u;  // groups started up to and including r, - 1
        if (b) ws.bnd[g] = r;
        if (r < r1w) gidx[r] = g;
        gpos += __popc(m);
    }
    if (t == 0) ws.bnd[U] = R;
    __syncthreads();

    // ---- F. size
    uint32_t loc = 0;
    for (uint32_t g = t; g < U; g += T) {
        uint32_t b = ws.bnd[g];
        double v = __longlong_as_double((long long)ks[b]);
        loc += value_bytes(v, bitdepth) + varint_len(ws.bnd[g + 1] - b);
    }
    uint32_t grp_bytes = block_sum_u32(loc, sh);
    uint32_t hdr = 2 + varint_len(U);
    uint32_t size = hdr + grp_bytes + fw->rle_idx_bytes;
    if (t == 0) {
        fw->rle_groups = U;
        fw->rle_size = size;
        fw->rle_valid = 1;
    }
    if (out == nullptr) {
        __syncthreads();
        return size;
    }

    // ---- G. emission: per run its start index; the first run of a group is preceded by the
    // group's value and run count (rle.rs:40-67)
    if (t == 0) {
        out[0] = 60;  // RLE_COMPRESSOR_ID (rle.rs:27)
        out[1] = (uint8_t)bitdepth;
        put_varint(out + 2, U);
    }
    auto item = [&](uint32_t r, bool &b, double &v, uint32_t &cntg, uint32_t &idx) -> uint32_t {
        if (r >= r1w) return 0u;
        idx = is[r];
        uint32_t len = varint_len(idx);
        b = r == 0 || ks[r] != ks[r - 1];
        if (b) {
            const uint32_t g = gidx[r];
            v = __longlong_as_double((long long)ks[r]);
            cntg = ws.bnd[g + 1] - ws.bnd[g];
            len += value_bytes(v, bitdepth) + varint_len(cntg);
        }
        return len;
    };
    uint32_t bytes = 0;
    for (uint32_t rb = r0w; rb < r1w; rb += 32) {
        bool b = false;
        double v = 0.0;
        uint32_t cntg = 0, idx = 0;
        bytes += item(rb + lane, b, v, cntg, idx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_down_sync(0xffffffffu, bytes, o);
    uint32_t tot_bytes;
    uint32_t base = hdr + rle_warp_offsets(bytes, sh, &tot_bytes);
    for (uint32_t rb = r0w; rb < r1w; rb += 32) {
        const uint32_t r = rb + lane;
        bool b = false;
        double v = 0.0;
        uint32_t cntg = 0, idx = 0;
        const uint32_t len = item(r, b, v, cntg, idx);
        uint32_t inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (r < r1w) {
            uint8_t *p = out + base + inc - len;
            if (b) {
                p += put_value(p, v, bitdepth);
                p += put_varint(p, cntg);
            }
            put_varint(p, idx);
        }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncthreads();
    return size;
}

}  // namespace atsc
