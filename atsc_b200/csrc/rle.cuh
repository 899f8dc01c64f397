// rle.cuh -- index RLE (rle.rs:142-189 IndexRLE::new, :40-67 Encode), device side.
// The reference's BTreeMap<u64 bits, Vec<usize>> is a group-by over run starts ordered by
// f64::to_bits; here: ordered compaction of runs, stable LSD radix sort by the 64-bit key
// (4-bit digits, uniform digits skipped), boundary detection, size / emission by prefix sums.
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

// All threads of the CTA call.  Returns the exact payload size; if out != nullptr also
// writes the payload.  sh: 128-word static scratch; hist: RLE_HIST_WORDS words.
__device__ inline uint32_t rle_process(const double *__restrict__ d, FrameWork *fw, RleWs ws,
                                       uint8_t *out, uint32_t *sh, uint32_t *hist) {
    const uint32_t N = fw->len;
    const int bitdepth = fw->bitdepth;
    const uint32_t T = blockDim.x, t = threadIdx.x;

    // ---- A. ordered compaction of run ends
    uint32_t R = 0;
    for (uint32_t i0 = 0; i0 < N; i0 += T) {
        uint32_t i = i0 + t;
        bool end = false;
        double v = 0.0;
        if (i < N) {
            v = d[i];
            end = (i + 1 >= N) || (d[i + 1] != v);
        }
        uint32_t tot;
        uint32_t r = block_excl_scan_u32(end ? 1u : 0u, sh, &tot);
        if (end) {
            ws.k0[R + r] = (uint64_t)__double_as_longlong(v);  // bits of the run's last element
            ws.bnd[R + r] = i;
        }
        R += tot;
        __syncthreads();
    }
    // ---- B. run starts
    for (uint32_t r = t; r < R; r += T) ws.i0[r] = r == 0 ? 0u : ws.bnd[r - 1] + 1u;
    // ---- C. which digits vary
    uint64_t key0 = ws.k0[0];
    uint32_t dlo = 0, dhi = 0;
    for (uint32_t r = t; r < R; r += T) {
        uint64_t x = ws.k0[r] ^ key0;
        dlo |= (uint32_t)x;
        dhi |= (uint32_t)(x >> 32);
    }
    // OR-reduce through shared memory atomics
    __syncthreads();
    if (t == 0) {
        sh[100] = 0;
        sh[101] = 0;
    }
    __syncthreads();
    if (dlo) atomicOr(&sh[100], dlo);
    if (dhi) atomicOr(&sh[101], dhi);
    __syncthreads();
    uint64_t diff = ((uint64_t)sh[101] << 32) | sh[100];
    __syncthreads();

    // ---- D. stable LSD radix sort on the varying 4-bit digits
    uint64_t *ks = ws.k0, *kd = ws.k1;
    uint32_t *is = ws.i0, *id = ws.i1;
    const uint32_t c = (R + T - 1) / T;
    const uint32_t lo = min(R, t * c), hi = min(R, (t + 1) * c);
    for (int p = 0; p < 16; p++) {
        if (((diff >> (4 * p)) & 15ull) == 0) continue;
        // column t of hist is private to thread t (bank = t mod 32: conflict free)
#pragma unroll
        for (int q = 0; q < 16; q++) hist[q * T + t] = 0;
        for (uint32_t r = lo; r < hi; r++) {
            uint32_t dg = (uint32_t)(ks[r] >> (4 * p)) & 15u;
            hist[dg * T + t]++;
        }
        uint32_t running = 0;
        for (int q = 0; q < 16; q++) {
            uint32_t tot;
            uint32_t ex = block_excl_scan_u32(hist[q * T + t], sh, &tot);
            hist[q * T + t] = running + ex;
            running += tot;
        }
        for (uint32_t r = lo; r < hi; r++) {
            uint64_t key = ks[r];
            uint32_t dg = (uint32_t)(key >> (4 * p)) & 15u;
            uint32_t dst = hist[dg * T + t]++;
            kd[dst] = key;
            id[dst] = is[r];
        }
        __syncthreads();
        uint64_t *tk = ks;
        ks = kd;
        kd = tk;
        uint32_t *ti = is;
        is = id;
        id = ti;
    }

    // ---- E. group boundaries (ordered compaction of positions where the key changes)
    uint32_t U = 0;
    for (uint32_t r0 = 0; r0 < R; r0 += T) {
        uint32_t r = r0 + t;
        bool b = r < R && (r == 0 || ks[r] != ks[r - 1]);
        uint32_t tot;
        uint32_t g = block_excl_scan_u32(b ? 1u : 0u, sh, &tot);
        if (b) ws.bnd[U + g] = r;
        U += tot;
        __syncthreads();
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

    // ---- G. emission
    if (t == 0) {
        out[0] = 60;  // RLE_COMPRESSOR_ID (rle.rs:27)
        out[1] = (uint8_t)bitdepth;
        put_varint(out + 2, U);
    }
    uint32_t base = hdr, gbase = 0;
    for (uint32_t r0 = 0; r0 < R; r0 += T) {
        uint32_t r = r0 + t;
        bool in = r < R;
        bool b = in && (r == 0 || ks[r] != ks[r - 1]);
        uint32_t gtot;
        uint32_t g = gbase + block_excl_scan_u32(b ? 1u : 0u, sh, &gtot);
        __syncthreads();
        uint32_t len = 0, cntg = 0;
        double v = 0.0;
        uint32_t idx = 0;
        if (in) {
            idx = is[r];
            len = varint_len(idx);
            if (b) {
                v = __longlong_as_double((long long)ks[r]);
                cntg = ws.bnd[g + 1] - ws.bnd[g];
                len += value_bytes(v, bitdepth) + varint_len(cntg);
            }
        }
        uint32_t tot;
        uint32_t off = block_excl_scan_u32(len, sh, &tot);
        if (in) {
            uint8_t *p = out + base + off;
            if (b) {
                p += put_value(p, v, bitdepth);
                p += put_varint(p, cntg);
            }
            put_varint(p, idx);
        }
        base += tot;
        gbase += gtot;
        __syncthreads();
    }
    return size;
}

}  // namespace atsc
