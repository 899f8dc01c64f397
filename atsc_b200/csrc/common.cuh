// common.cuh -- shared device structs and helpers for the ATSC sm_100a kernels.
//
// Reference citations are relative to /root/reference/atsc/src/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace atsc {

// Bitdepth variant index (optimizer/utils.rs:20-26)
enum { BD_F64 = 0, BD_I32 = 1, BD_I16 = 2, BD_U8 = 3 };
// Compressor variant index (compressor/mod.rs:34-44)
enum { C_NOOP = 0, C_FFT = 1, C_IDW = 2, C_CONSTANT = 3, C_POLY = 4, C_AUTO = 5, C_RLE = 6 };

constexpr int MAX_FRAME = 131072;      // optimizer/mod.rs:27 MAX_FRAME_SIZE
constexpr int MAX_FFT_LEN = 139968;    // next_size(131072) = 2^6 * 3^7
constexpr int BLOCK = 1024;            // threads per CTA for the per-frame kernels

constexpr uint8_t TIE_FFT_LOOP = 1, TIE_POLY_LOOP = 2, TIE_SELECT = 4, TIE_FFT_TOPK = 8;

// FrameWork.front_mode: the frame goes through k_front (stats in the streaming pass); it also evaluates
// the first Polynomial step there; it also accumulates the FFT probe's stage-1 fold there
constexpr uint8_t FM_ON = 1, FM_POLY = 2, FM_FOLD = 4;
// ... or through k_sfold (sfold.cuh): its stats partials are per slot-range item and its probe fold sits in the fold arena
constexpr uint8_t FM_SFOLD = 8;
// FrameWork.front_res: poly_step / poly_err hold the first step's error (the loop did not end there)
constexpr uint8_t FRES_POLY1 = 1;
// ... k_probe found enough nonzero bins and the FFT candidate is NOT pruned: k_fft_fwd goes straight to the transform
constexpr uint8_t FRES_SURVIVOR = 4;

// k_poly1s / k_poly1 (poly.cuh): work items of the first Polynomial step of the big frames
constexpr uint32_t POLY_ITEM = 32768;          // samples per item
constexpr uint32_t POLY_ITEM_MIN_LEN = 65536;  // frames this long go through k_poly1s (step 100, >= 2 items)
__host__ __device__ inline uint32_t poly_item_count(uint32_t N) { return (N + POLY_ITEM - 1) / POLY_ITEM; }
// the step Polynomial::compress_bounded tries first
__host__ __device__ inline uint32_t poly_first_step(uint32_t N) {
    const uint32_t baseline = (3 >= N / 100) ? 3 : N / 100;  // polynomial.rs:218-221
    const uint32_t step = N / baseline;
    return step < 1 ? 1 : step;
}
constexpr uint32_t POLY_ITEM_KEYS = 352;       // keys / tangents of one item (<= 82 blocks of 4 segments + the tail)
// Part q of the Q = poly_item_count(N) parts of a frame's first step (step 100): the four-segment blocks [b_lo, b_hi) of the
// nblk = (K - 3) / 4 whole blocks that cover the Catmull-Rom segments 1 .. 4 * nblk (K keys: polynomial.rs:329-340); the
// segments 4 * nblk + 1 .. K - 3 and the Linear ends are poly_first_step_rest's.
__host__ __device__ inline void p1_item_blocks(uint32_t N, uint32_t q, uint32_t *b_lo, uint32_t *b_hi) {
    const uint32_t kreg = (N + 99u) / 100u, K = kreg + (((kreg - 1u) * 100u != N - 1u) ? 1u : 0u);
    const uint32_t Q = poly_item_count(N), nblk = (K - 3u) / 4u;
    *b_lo = (uint32_t)((uint64_t)nblk * q / Q);
    *b_hi = (uint32_t)((uint64_t)nblk * (q + 1u) / Q);
}
// k_poly1s (poly.cuh): self-contained descriptors of its work items, appended by k_plan
constexpr uint32_t P1_STEP = 100;
constexpr int P1_T = 512;                           // threads of k_poly1s: five groups of 100
constexpr uint32_t P1_G = P1_T / P1_STEP;           // groups
constexpr uint32_t P1_PARTS = P1_T / 32;            // partial sums per item: one per warp
struct alignas(16) P1Item {
    const double *d;       // the frame's samples
    double vmin, vmax;
    uint32_t N;            // frame length
    uint32_t b_lo, b_hi;   // NS-blocks of the item: segments 1 + NS * b_lo .. NS * b_hi
    uint32_t out;          // parts[out * P1_PARTS + warp]
    uint32_t nkeys;        // keys 1 + NS * b_lo .. NS * b_hi + 1 of the item
    uint32_t tame, pad[4];
};
static_assert(sizeof(P1Item) == 64, "P1Item is copied as four 16-byte pieces");

// One record per frame, lives in device memory for the duration of a wave.
struct FrameWork {
    // ---- input
    uint64_t off;  // first sample in the device sample buffer
    uint32_t len;  // N
    uint8_t comp;  // requested compressor
    uint8_t bounded;
    uint8_t select_only;  // sampled-selection pass (frame/mod.rs:94-105): only the winner is wanted
    uint8_t forced;  // Auto frames only: 0xFF = pick freely, else the compressor the sampled pass chose
    // ---- stats (optimizer/utils.rs DataStats)
    double vmin, vmax;
    uint8_t bitdepth, fractional, is_const, f32_const;
    uint32_t n_runs;
    uint32_t rle_idx_bytes;
    // ---- which candidates run
    uint8_t need_poly, need_rle, need_fft, poly_type;  // poly_type: 0 Catmull-Rom, 1 IDW
    // ---- polynomial / idw candidate
    uint32_t poly_step, poly_npts, poly_size;
    uint16_t poly_iters;
    uint8_t poly_tie, poly_valid;
    double poly_err;
    // ---- rle candidate
    uint32_t rle_groups, rle_size;
    uint8_t rle_valid;
    uint8_t fft_small;  // transform length <= 1152: k_fft_small (fft_small.cuh) runs the FFT candidate
    uint8_t front_mode;  // FM_* bits: what k_front (front.cuh) does for this frame (set by the host)
    uint8_t front_res;   // FRES_* bits: what k_front left behind
    // ---- fft candidate
    uint32_t fft_count, fft_size;
    uint16_t fft_iters;
    uint8_t fft_tie, fft_valid;  // fft_valid: 1 evaluated to the reference's stopping point, 2 pruned (cannot win)
    double fft_err;
    uint64_t fft_list_off;  // entry offset into the wave's FftEntry arena
    uint32_t fft_list_cap;
    int32_t geom;        // index into the FftGeom table, -1 = direct DFT (N < 128)
    uint32_t aux_size;   // Noop / Constant payload size
    uint32_t fwd_done;   // k_fft_fwd left the half spectrum + keys at spec_off
    uint32_t chunk0;     // first entry of this frame in the wave's stats chunk table
    uint32_t fold_slot;  // FM_SFOLD: this frame's block of the wave's fold arena
    uint32_t poly_part0, poly_parts;  // k_poly1s / k_poly1: first item / number of items of this frame's first-step partial sums (0: none)
    uint64_t spec_off;   // entry offset into the wave's spectrum arena (~0 = frame not eligible for fft2.cuh)
    // ---- result
    uint8_t winner, near_tie;
    uint16_t iterations;
    uint32_t payload_len;
    uint64_t payload_off;
    double error;
};

// sorted (descending |z|) spectrum entry kept for emission
struct FftEntry {
    uint32_t bin;  // true bin index 0..L/2 (u16 wrap applied at use, fft.rs:242)
    float re, im;
};

// ----------------------------------------------------------------------------
// bincode varint helpers (compressor/mod.rs:126-130, bincode 2 "standard")
// ----------------------------------------------------------------------------
__host__ __device__ inline uint32_t varint_len(uint64_t u) {
    return u < 251 ? 1u : u < 65536ull ? 3u : u < 4294967296ull ? 5u : 9u;
}
__host__ __device__ inline uint64_t zigzag64(int64_t n) {
    return ((uint64_t)n << 1) ^ (uint64_t)(n >> 63);
}
__host__ __device__ inline int64_t unzigzag64(uint64_t u) {
    return (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
}
__host__ __device__ inline uint32_t put_varint(uint8_t *p, uint64_t u) {
    if (u < 251) {
        p[0] = (uint8_t)u;
        return 1;
    }
    if (u < 65536ull) {
        p[0] = 251;
        p[1] = (uint8_t)u;
        p[2] = (uint8_t)(u >> 8);
        return 3;
    }
    if (u < 4294967296ull) {
        p[0] = 252;
        for (int i = 0; i < 4; i++) p[1 + i] = (uint8_t)(u >> (8 * i));
        return 5;
    }
    p[0] = 253;
    for (int i = 0; i < 8; i++) p[1 + i] = (uint8_t)(u >> (8 * i));
    return 9;
}
// length of the varint that starts with byte b (0 = invalid marker)
__host__ __device__ inline uint32_t varint_len_from_first(uint8_t b) {
    return b < 251 ? 1u : b == 251 ? 3u : b == 252 ? 5u : b == 253 ? 9u : 17u;
}
__host__ __device__ inline uint64_t get_varint(const uint8_t *p, uint32_t *len) {
    uint8_t b = p[0];
    if (b < 251) {
        *len = 1;
        return b;
    }
    int nb = b == 251 ? 2 : b == 252 ? 4 : 8;
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v |= (uint64_t)p[1 + i] << (8 * i);
    *len = 1 + nb;
    return v;
}
__host__ __device__ inline void put_bytes(uint8_t *p, const void *src, int n) {
    const uint8_t *s = (const uint8_t *)src;
    for (int i = 0; i < n; i++) p[i] = s[i];
}

// ----------------------------------------------------------------------------
// Rust `as` casts: saturating, NaN -> 0
// ----------------------------------------------------------------------------
__device__ inline int64_t rust_as_i64(double x) {
    return __double2ll_rz(x);  // cvt.rzi.s64.f64 saturates, NaN -> 0x8000.. on some archs
}
__device__ inline int32_t rust_as_i32(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int32_t)0x80000000;
    return (int32_t)x;
}
__device__ inline int32_t rust_as_i16(double x) {
    if (x != x) return 0;
    if (x >= 32767.0) return 32767;
    if (x <= -32768.0) return -32768;
    return (int32_t)x;
}
__device__ inline uint32_t rust_as_u8(double x) {
    if (x != x) return 0;
    if (x >= 255.0) return 255;
    if (x <= 0.0) return 0;
    return (uint32_t)x;
}
__device__ inline int64_t rust_as_i64_safe(double x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0) return 0x7FFFFFFFFFFFFFFFll;
    if (x <= -9223372036854775808.0) return (int64_t)0x8000000000000000ull;
    return (int64_t)x;
}

// number of payload bytes one stored value takes at a given bitdepth
// (polynomial.rs:61-81, rle.rs:47-64, constant.rs:44-61)
__device__ inline uint32_t value_bytes(double v, int bitdepth) {
    switch (bitdepth) {
        case BD_U8: return 1;
        case BD_I16: return varint_len(zigzag64((int64_t)rust_as_i16(v)));
        case BD_I32: return varint_len(zigzag64((int64_t)rust_as_i32(v)));
        default: return 8;
    }
}
__device__ inline uint32_t put_value(uint8_t *p, double v, int bitdepth) {
    switch (bitdepth) {
        case BD_U8: p[0] = (uint8_t)rust_as_u8(v); return 1;
        case BD_I16: return put_varint(p, zigzag64((int64_t)rust_as_i16(v)));
        case BD_I32: return put_varint(p, zigzag64((int64_t)rust_as_i32(v)));
        default: put_bytes(p, &v, 8); return 8;
    }
}

// ----------------------------------------------------------------------------
// utils/mod.rs:61-74 rounding.  No FMA contraction: explicit _rn intrinsics.
// ----------------------------------------------------------------------------
// round half away from zero (Rust f64::round == C round)
__device__ inline double round5_exact(double x) {
    // (x * 1e5).round() / 1e5 with a true IEEE division
    return __ddiv_rn(round(__dmul_rn(x, 100000.0)), 100000.0);
}
__device__ inline double round_and_limit5(double x, double mn, double mx) {
    double out = round5_exact(x);
    if (out < mn) return mn;
    if (out > mx) return mx;
    return out;
}
__device__ inline double round_f64_dec(double x, int decimals) {
    double y = decimals == 3 ? 1000.0 : decimals == 4 ? 10000.0 : 100000.0;
    return __ddiv_rn(round(__dmul_rn(x, y)), y);
}

// f64::round (half away from zero) without the library call: trunc(y + copysign(pred(0.5), y)).
// Exact for every double: a fraction below one half stays below the next integer after the
// addition (the sum is at least 2^-54 short of it, which survives rounding at any exponent where
// fractions exist), a fraction of one half or more reaches it, and from 2^52 on y is an integer
// that the addend cannot move.  NaN, +-inf and -0.0 pass through like f64::round.
__device__ __forceinline__ double round_half_away(double y) {
    return trunc(__dadd_rn(y, copysign(0.49999999999999994, y)));
}
// n / 100000.0, correctly rounded, for the integer-valued n that `round()` returns: Markstein's
// two-step FMA refinement of n * RN(1/100000) (q1 is already faithful, q2 is the IEEE quotient;
// checked against the hardware division on 4e8 random integers of every width up to 53 bits).
// Branch free: +-inf (where the refinement would produce NaN) is passed through by a select.
__device__ __forceinline__ double div_1e5(double n) {
    const double b = 100000.0, y = 1e-5;
    double q = __dmul_rn(n, y);
    q = __fma_rn(__fma_rn(-b, q, n), y, q);
    q = __fma_rn(__fma_rn(-b, q, n), y, q);
    return fabs(n) == __longlong_as_double(0x7FF0000000000000ll) ? n : q;
}
// The same quotient for an INTEGER n with |n| < 2^53 in one refinement step.  q0 = RN(n*y) is within
// 1.24 ulp of n/1e5 (y = RN(1e-5) is 2^-53.44 relative above 1e-5), so r = n - 1e5*q0 is a multiple
// of 32 ulp(q0) below 2^12 of them: exact.  q0 + r/1e5 IS n/1e5, r*y differs from r/1e5 by < 1e-16
// ulp, and n/1e5 is either a double or at least 5e-6 ulp away from every midpoint between doubles
// (it cannot be a midpoint: 3125 | n would make it a 42-bit dyadic), so the last rounding lands on
// RN(n/1e5).  No inf pass-through: callers guarantee finite n (tools/div_check.py replays the proof
// in exact rational arithmetic).
__device__ __forceinline__ double div_1e5_int53(double n) {
    const double b = 100000.0, y = 1e-5;
    const double q = __dmul_rn(n, y);
    return __fma_rn(__fma_rn(-b, q, n), y, q);
}
// 1 / o to ~2^-45 relative, branch free: hardware seed (2^-23) + one Newton step.  The error loops
// only need the MAPE terms far inside the near-tie margins (1e-10 absolute on the mean, see
// mape_term).  o == 0 or denormal -> the seed itself (+-inf), like the IEEE quotient of a zero
// sample; inf / NaN propagate through the seed as well.
__device__ __forceinline__ double rcp_fast(double o) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(o));
    // seed +-inf (o zero / denormal), 0 (o = +-inf) or NaN all turn the Newton step into NaN: keep the seed
    const double y1 = __fma_rn(y, __fma_rn(-o, y, 1.0), y);
    return y1 == y1 ? y1 : y;
}
// rcp_fast for a normal, finite, nonzero o well inside the exponent range: no NaN select needed
__device__ __forceinline__ double rcp_fast_tame(double o) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(o));
    return __fma_rn(y, __fma_rn(-o, y, 1.0), y);
}
// One MAPE term |(out - o) / o| (utils/error.rs:110-113) through a reciprocal good to ~3e-14
// relative (absorbed by the near-tie tolerance, the error only feeds threshold tests); an exactly
// reproduced sample gives exactly 0.  A zero sample keeps the reference's
// semantics: w = inf gives inf (out != 0) or NaN (out == 0), SURVEY H5.
__device__ __forceinline__ double mape_term(double out, double o) {
    return fabs(__dmul_rn(__dsub_rn(out, o), rcp_fast(o)));
}
__device__ __forceinline__ double mape_term_tame(double out, double o) {
    return fabs(__dmul_rn(__dsub_rn(out, o), rcp_fast_tame(o)));
}

// optimizer/utils.rs:115-160 split_n: (integer part as i64, fraction != 0)
__device__ inline int64_t split_n(double x, bool *frac_nz) {
    uint64_t bits = (uint64_t)__double_as_longlong(x);
    bool neg = ((int64_t)bits) < 0;
    int exponent = (int)((bits >> 52) & 0x7FF);
    uint64_t m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    int64_t mant = neg ? -(int64_t)m : (int64_t)m;
    int shl = exponent + (64 - 53 - 1023 + 1);
    if (shl <= 0) {
        int shr = -shl;
        if (shr < 64) {
            *frac_nz = (((uint64_t)mant) >> shr) != 0;
            return 0;
        }
        *frac_nz = false;
        return 0;
    } else if (shl < 64) {
        *frac_nz = (((uint64_t)mant) << shl) != 0;
        return mant >> (64 - shl);
    } else if (shl < 128) {
        *frac_nz = false;
        return (int64_t)(((uint64_t)mant) << (shl - 64));
    }
    *frac_nz = false;
    return 0;
}
// optimizer/utils.rs:91-113
__device__ inline int bitdepth_of(int64_t max_int, int64_t min_int) {
    int bd = max_int <= 255 ? 8 : max_int <= 32767 ? 16 : max_int <= 2147483647ll ? 32 : 64;
    int bs = (min_int >= 0 && min_int <= 255) ? 8
             : min_int >= -32768             ? 16
             : min_int >= -2147483648ll      ? 32
                                             : 64;
    int b = bd > bs ? bd : bs;
    return b == 8 ? BD_U8 : b == 16 ? BD_I16 : b == 32 ? BD_I32 : BD_F64;
}

// ----------------------------------------------------------------------------
// block-wide primitives (BLOCK threads, warp shuffles + one smem hop)
// ----------------------------------------------------------------------------
__device__ inline double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// deterministic tree sum; result valid in every thread. scratch: >= 33 doubles
__device__ inline double block_sum(double v, double *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ inline uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// scratch: >= 33 uint32
__device__ inline uint32_t block_sum_u32(uint32_t v, uint32_t *scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum_u32(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < (blockDim.x >> 5) ? scratch[lane] : 0u;
        t = warp_sum_u32(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
// exclusive scan of one uint32 per thread; returns exclusive prefix, *total = block sum.
// scratch: >= 33 uint32
__device__ inline uint32_t block_excl_scan_u32(uint32_t v, uint32_t *scratch, uint32_t *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
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
            if (lane >= o) ti += u;
        }
        scratch[lane] = ti - t;  // exclusive warp offsets
        if (lane == 31) scratch[32] = ti;
    }
    __syncthreads();
    uint32_t res = scratch[w] + inc - v;
    *total = scratch[32];
    return res;
}

// dynamic work queue: every CTA pulls the next item index
__device__ inline int queue_next(unsigned int *counter, int *smem_slot) {
    __syncthreads();
    if (threadIdx.x == 0) *smem_slot = (int)atomicAdd(counter, 1u);
    __syncthreads();
    return *smem_slot;
}

}  // namespace atsc
