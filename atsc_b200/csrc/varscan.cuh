// varscan.cuh -- CTA-parallel walk of a byte stream of variable-length records whose
// length is decided by their first byte (bincode varints, compressor/mod.rs:126-130;
// FFT FrequencyPoint = varint(u16) + 2 x f32, fft.rs:33-39).
//
// Record boundaries form a dependent chain, so the stream is cut into 32-byte chunks (one per
// thread); for every possible entry offset a thread computes where its chunk hands over to
// the next one, the hand-over maps are composed with a block-wide scan (function composition
// is associative), and then every thread knows its true entry offset and first record index.
#pragma once
#include "common.cuh"

namespace atsc {

constexpr int VS_CHUNK = 32;

// maps are packed 4 bits per entry state (states 0..MAXLEN-1, MAXLEN <= 12)
template <int NS>
__device__ inline unsigned long long vs_compose(unsigned long long first, unsigned long long then) {
    unsigned long long r = 0;
#pragma unroll
    for (int e = 0; e < NS; e++) {
        unsigned a = (unsigned)(first >> (4 * e)) & 15u;
        unsigned b = (unsigned)(then >> (4 * a)) & 15u;
        r |= (unsigned long long)b << (4 * e);
    }
    return r;
}
template <int NS>
__device__ inline unsigned long long vs_identity() {
    unsigned long long r = 0;
#pragma unroll
    for (int e = 0; e < NS; e++) r |= (unsigned long long)e << (4 * e);
    return r;
}

struct VarintLen {
    static constexpr int MAXLEN = 9;
    __device__ static inline uint32_t len(uint8_t b) { return b < 251 ? 1u : b == 251 ? 3u : b == 252 ? 5u : 9u; }
};
struct FftEntryLen {
    static constexpr int MAXLEN = 11;
    __device__ static inline uint32_t len(uint8_t b) { return b < 251 ? 9u : 11u; }
};

// Walks `n_rec` records starting at bytes[0] (at most nbytes available).  For every record r,
// calls fn(r, offset).  Returns the total number of bytes the n_rec records occupy
// (0xFFFFFFFF if the stream is too short).  All threads of the CTA must call.
// sh: 128 words; sm64: >= blockDim.x/32 + 2 u64 of shared scratch.
template <class LF, class Fn>
__device__ inline uint32_t varscan(const uint8_t *__restrict__ bytes, uint32_t nbytes, uint32_t n_rec,
                                   uint32_t *sh, unsigned long long *sm64, Fn fn) {
    constexpr int NS = LF::MAXLEN;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const int lane = t & 31, w = t >> 5, nw = T >> 5;
    uint32_t state = 0;      // entry offset of the tile's first chunk
    uint32_t rec_base = 0;   // records before this tile
    uint32_t end_off = 0xFFFFFFFFu;
    bool done = n_rec == 0;
    if (done) end_off = 0;
    for (uint32_t tile0 = 0; !done; tile0 += T * VS_CHUNK) {
        const uint32_t c0 = tile0 + t * VS_CHUNK;  // first byte of my chunk
        // hand-over map for every entry state
        unsigned long long map = 0;
#pragma unroll 1
        for (int e = 0; e < NS; e++) {
            uint32_t o = c0 + e;
            const uint32_t lim = c0 + VS_CHUNK;
            while (o < lim && o < nbytes) o += LF::len(bytes[o]);
            uint32_t ex = o >= lim ? o - lim : 0;  // (stream end: value irrelevant)
            if (ex > (uint32_t)(NS - 1)) ex = NS - 1;
            map |= (unsigned long long)ex << (4 * e);
        }
        // inclusive scan of compositions
        unsigned long long inc = map;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = vs_compose<NS>(prev, inc);
        }
        __syncthreads();
        if (lane == 31) sm64[w] = inc;
        __syncthreads();
        if (w == 0) {
            unsigned long long v = lane < nw ? sm64[lane] : vs_identity<NS>();
            unsigned long long vi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long prev = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) vi = vs_compose<NS>(prev, vi);
            }
            // exclusive: composition of all previous warps
            unsigned long long ex = __shfl_up_sync(0xffffffffu, vi, 1);
            if (lane == 0) ex = vs_identity<NS>();
            sm64[lane] = ex;
            if (lane == 31) sm64[32] = vi;  // whole tile
        }
        __syncthreads();
        unsigned long long before_warp = sm64[w];
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = vs_identity<NS>();
        unsigned long long upto_me = vs_compose<NS>(before_warp, excl);  // all chunks before mine
        uint32_t my_state = (uint32_t)(upto_me >> (4 * state)) & 15u;
        uint32_t next_state = (uint32_t)(sm64[32] >> (4 * state)) & 15u;
        // count records that start in my chunk
        uint32_t cnt = 0;
        {
            uint32_t o = c0 + my_state;
            const uint32_t lim = c0 + VS_CHUNK;
            while (o < lim && o < nbytes) {
                o += LF::len(bytes[o]);
                cnt++;
            }
        }
        uint32_t tot;
        uint32_t first = rec_base + block_excl_scan_u32(cnt, sh, &tot);
        // visit
        {
            uint32_t o = c0 + my_state;
            const uint32_t lim = c0 + VS_CHUNK;
            uint32_t r = first;
            while (o < lim && o < nbytes && r < n_rec) {
                uint32_t l = LF::len(bytes[o]);
                if (o + l <= nbytes) fn(r, o);
                if (r == n_rec - 1) end_off = o + l;
                o += l;
                r++;
            }
        }
        rec_base += tot;
        state = next_state;
        if (rec_base >= n_rec || tile0 + T * VS_CHUNK >= nbytes) done = true;
        __syncthreads();
    }
    // broadcast end offset (exactly one thread saw the last record)
    __syncthreads();
    if (t == 0) sh[104] = 0xFFFFFFFFu;
    __syncthreads();
    if (end_off != 0xFFFFFFFFu && (n_rec == 0 ? t == 0 : true)) atomicMin(&sh[104], end_off);
    __syncthreads();
    uint32_t res = sh[104];
    __syncthreads();
    if (res != 0xFFFFFFFFu && res > nbytes) res = 0xFFFFFFFFu;
    return res;
}

// inclusive max-scan of one uint32 per thread (used by the RLE fill); scratch >= 33 words
__device__ inline uint32_t block_incl_scan_max_u32(uint32_t v, uint32_t *scratch, uint32_t *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    __syncthreads();
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < (blockDim.x >> 5) ? scratch[lane] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti = max(ti, u);
        }
        uint32_t ex = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane == 0) ex = 0;
        scratch[lane] = ex;
        if (lane == 31) scratch[32] = ti;
    }
    __syncthreads();
    uint32_t res = max(scratch[w], inc);
    *total = scratch[32];
    return res;
}

}  // namespace atsc
