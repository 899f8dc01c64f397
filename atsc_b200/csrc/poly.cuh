// poly.cuh -- Polynomial (Catmull-Rom) and IDW compressors, device side.
//   Polynomial::compress_bounded   polynomial.rs:209-277
//   Polynomial::compress_hinted    polynomial.rs:279-305
//   Polynomial::polynomial_to_data polynomial.rs:342-373  (splines 4.3.1 semantics)
//   Polynomial::idw_to_data        polynomial.rs:375-393  (inverse_distance_weight 0.1.1)
// All value arithmetic uses explicit _rn intrinsics: identical operation order to the
// reference and no FMA contraction, so decompressed values are bit-exact.
#pragma once
#include "common.cuh"

namespace atsc {

// key layout of one candidate step (polynomial.rs:329-340 get_positions)
struct PolyKeys {
    uint32_t N, step, Kreg, K;  // Kreg = ceil(N/step) regular keys; K = Kreg (+1 if last index appended)
};
__host__ __device__ inline PolyKeys poly_keys(uint32_t N, uint32_t step) {
    PolyKeys k;
    k.N = N;
    k.step = step;
    k.Kreg = (N + step - 1) / step;
    k.K = k.Kreg + (((k.Kreg - 1) * step != N - 1) ? 1u : 0u);
    return k;
}
__device__ inline uint32_t poly_pos(const PolyKeys &k, uint32_t j) {
    return j < k.Kreg ? j * k.step : k.N - 1;
}

// splines 4.3.1 cubic_hermite with (t, value) pairs x(before a), a, b, y(after b)
__device__ inline double cubic_hermite(double t, double xt, double xv, double at, double av,
                                       double bt, double bv, double yt, double yv) {
    double two_t = __dmul_rn(t, 2.0);
    double three_t = __dmul_rn(t, 3.0);
    double t2 = __dmul_rn(t, t);
    double t3 = __dmul_rn(t2, t);
    double two_t3 = __dmul_rn(t2, two_t);
    double two_t2 = __dmul_rn(t, two_t);
    double three_t2 = __dmul_rn(t, three_t);
    double seg = __dsub_rn(bt, at);
    double m0 = __dmul_rn(__ddiv_rn(__dsub_rn(bv, xv), __dsub_rn(bt, xt)), seg);
    double m1 = __dmul_rn(__ddiv_rn(__dsub_rn(yv, av), __dsub_rn(yt, at)), seg);
    double c0 = __dmul_rn(av, __dadd_rn(__dsub_rn(two_t3, three_t2), 1.0));
    double c1 = __dmul_rn(m0, __dadd_rn(__dsub_rn(t3, two_t2), t));
    double c2 = __dmul_rn(bv, __dsub_rn(three_t2, two_t3));
    double c3 = __dmul_rn(m1, __dsub_rn(t3, t2));
    return __dadd_rn(__dadd_rn(__dadd_rn(c0, c1), c2), c3);
}

// Spline value at integer x.  `pts` yields the value of key j.
template <class PtsFn>
__device__ inline double poly_eval_at(const PolyKeys &k, uint32_t x, PtsFn pts) {
    if (k.K == 1) return pts(0);
    uint32_t lastpos = poly_pos(k, k.K - 1);
    if (x >= lastpos) return pts(k.K - 1);  // clamped_sample: t >= last.t -> last.value
    uint32_t i = x / k.step;                  // key i: pos[i] <= x < pos[i+1]
    double at = (double)poly_pos(k, i), bt = (double)poly_pos(k, i + 1);
    double nt = __ddiv_rn(__dsub_rn((double)x, at), __dsub_rn(bt, at));
    bool catmull = (i > 0) && (k.K - i > 2);  // polynomial.rs:349
    double av = pts(i), bv = pts(i + 1);
    if (!catmull) {
        // Linear: a * (1 - t) + b * t
        return __dadd_rn(__dmul_rn(av, __dsub_rn(1.0, nt)), __dmul_rn(bv, nt));
    }
    return cubic_hermite(nt, (double)poly_pos(k, i - 1), pts(i - 1), at, av, bt, bv,
                         (double)poly_pos(k, i + 2), pts(i + 2));
}

// IDW value at integer x: all K keys, power 2, sums in key order.
// inv_d2[d] = 1 / (d*d) (IEEE), d >= 1.
template <class PtsFn>
__device__ inline double idw_eval_at(const PolyKeys &k, uint32_t x, PtsFn pts,
                                     const double *__restrict__ inv_d2) {
    // exact hit -> that key's value
    if (x == k.N - 1) return pts(k.K - 1);
    if (x % k.step == 0 && x / k.step < k.Kreg) return pts(x / k.step);
    double S = 0.0;
    for (uint32_t j = 0; j < k.K; j++) {
        uint32_t p = poly_pos(k, j);
        uint32_t d = p > x ? p - x : x - p;
        S = __dadd_rn(S, inv_d2[d]);
    }
    // w_j / S for every key: the IEEE quotient through Markstein's two-step FMA refinement of
    // w * RN(1/S) (one true division per sample instead of one per key; q1 is faithful, q2 correctly
    // rounded -- checked against the hardware division on 6e8 (1/d^2, S) pairs).  The theorem's one
    // exception, a divisor whose mantissa is all ones, and non-finite sums take the plain division.
    const double y = __ddiv_rn(1.0, S);
    const unsigned long long sb = (unsigned long long)__double_as_longlong(S);
    const bool plain = (sb & 0x000FFFFFFFFFFFFFull) == 0x000FFFFFFFFFFFFFull || !(fabs(S) < __longlong_as_double(0x7FF0000000000000ll)) ||
                       !(fabs(y) < __longlong_as_double(0x7FF0000000000000ll));
    double acc = 0.0;
    for (uint32_t j = 0; j < k.K; j++) {
        uint32_t p = poly_pos(k, j);
        uint32_t d = p > x ? p - x : x - p;
        const double w = inv_d2[d];
        double q = __dmul_rn(w, y);
        q = __fma_rn(__fma_rn(-S, q, w), y, q);
        q = __fma_rn(__fma_rn(-S, q, w), y, q);
        if (plain) q = __ddiv_rn(w, S);
        acc = __dadd_rn(acc, __dmul_rn(q, pts(j)));
    }
    return acc;
}

// per-CTA-slot scratch of the refinement loop
struct PolyWs {
    double *slope;  // [MAX_FRAME + 8] tangent (central-difference slope * step) at every key of the current step
};
constexpr int POLY_MAXSTEP = 136;  // step <= 133 (polynomial.rs:290 with >= max(3, N/100) points)

// utils/mod.rs:66-74 round_and_limit_f64(x, min, max, 5), same value as round_and_limit5
__device__ __forceinline__ double round_and_limit5_fast(double x, double mn, double mx) {
    double n = round_half_away(__dmul_rn(x, 100000.0));
    double out = div_1e5(n);
    if (out < mn) return mn;
    if (out > mx) return mx;
    return out;
}

// The same for a "tame" frame (poly_tame): |x| * 1e5 stays an integer below 2^53 after rounding
__device__ __forceinline__ double round_and_limit5_tame(double x, double mn, double mx) {
    double out = div_1e5_int53(round_half_away(__dmul_rn(x, 100000.0)));
    if (out < mn) return mn;
    if (out > mx) return mx;
    return out;
}
// A frame is tame when every sample has the same sign and 1e-200 <= |sample| <= 1e10: no zero /
// denormal / non-finite denominators (rcp_fast_tame) and, a Catmull-Rom value staying within
// max|v| + 4/27 * (max - min) <= 1.3e10, no rounded value * 1e5 beyond 2^53 (div_1e5_int53).
// NaN samples compare false everywhere and come out as NaN on either path.
__device__ __forceinline__ bool poly_tame(double vmin, double vmax) {
    return (vmin >= 1e-200 || vmax <= -1e-200) && fmax(fabs(vmin), fabs(vmax)) <= 1e10;
}

// The part of a Catmull-Rom step's MAPE sum that the segment loop does not cover: the Linear ends
// (segment 0, segment K-2 -- possibly irregular -- and the last sample), through the generic
// per-sample arithmetic.  Returns this thread's share of the sum; all threads call.
__device__ inline double poly_mape_ends(const double *__restrict__ d, const PolyKeys &k, double vmin, double vmax) {
    const uint32_t N = k.N, step = k.step, K = k.K, T = blockDim.x, t = threadIdx.x;
    const double stepd = (double)step;
    auto pts = [&](uint32_t j) { return d[poly_pos(k, j)]; };
    double acc = 0.0;
    const uint32_t Kreg = k.Kreg;
    const uint32_t last_reg = (Kreg - 1) * step;  // position of the last regular key
    const uint32_t start_last = (K >= 4) ? (K - 2) * step : 0u;  // K < 4: no Catmull-Rom segment at all
    const uint32_t nA = min(step, start_last), nB = N - start_last;
    for (uint32_t e = t; e < nA + nB; e += T) {
        const uint32_t x = e < nA ? e : start_last + (e - nA);
        const uint32_t i = x / step, jj = x - i * step;
        double v;
        if (x == N - 1) {
            v = d[N - 1];
        } else if (i == Kreg - 1) {
            // irregular last segment [last_reg, N-1]: always Linear (it is segment K-2)
            double at = (double)last_reg, bt = (double)(N - 1);
            double nt = __ddiv_rn(__dsub_rn((double)x, at), __dsub_rn(bt, at));
            v = __dadd_rn(__dmul_rn(d[last_reg], __dsub_rn(1.0, nt)), __dmul_rn(d[N - 1], nt));
        } else if (i >= 1 && i + 2 < K) {
            // only reached when K < 4 cannot happen (then no such i exists); kept for completeness
            v = poly_eval_at(k, x, pts);
        } else {
            // first segment, or the regular segment K-2: Linear  a * (1 - t) + b * t
            const uint32_t pb = (i + 1 < Kreg) ? (i + 1) * step : N - 1;
            const double nt = __ddiv_rn((double)jj, stepd);
            v = __dadd_rn(__dmul_rn(d[i * step], __dsub_rn(1.0, nt)), __dmul_rn(d[pb], nt));
        }
        acc += mape_term(round_and_limit5_fast(v, vmin, vmax), d[x]);
    }
    return acc;
}

// This thread's share of a Catmull-Rom step's MAPE sum over the segments of the NS-blocks [b_lo, b_hi)
// (block b = segments 1 + NS*b .. NS*b + NS), plus -- `tail` -- the fewer-than-NS segments left over
// behind the last whole block.  tang[j - tbase] is the tangent of key j (global or shared memory).
// Thread t owns one offset j inside the segments (its Hermite basis values stay in registers) and
// walks over segments, so the inner loop has no table lookups, no index division and no branches.
constexpr int POLY_NS = 4;
template <bool TAME>
__device__ inline double poly_cr_range(const double *__restrict__ d, const PolyKeys &k, double vmin, double vmax,
                                       const double *__restrict__ tang, uint32_t tbase, uint32_t b_lo, uint32_t b_hi, bool tail) {
    const uint32_t step = k.step, K = k.K, T = blockDim.x, t = threadIdx.x;
    const double stepd = (double)step;
    double acc = 0.0;
    tang -= tbase;
    // Catmull-Rom segments 1 .. K-3 (polynomial.rs:349: key i is CatmullRom iff 0 < i < K-2)
    const uint32_t G = T / step, g = t / step, j = t - g * step;  // G groups of `step` threads
    if (g < G && K >= 4) {
        const double tt = __ddiv_rn((double)j, stepd);
        const double two_t = __dmul_rn(tt, 2.0), three_t = __dmul_rn(tt, 3.0);
        const double t2 = __dmul_rn(tt, tt), t3 = __dmul_rn(t2, tt);
        const double two_t3 = __dmul_rn(t2, two_t), two_t2 = __dmul_rn(tt, two_t), three_t2 = __dmul_rn(tt, three_t);
        const double h00 = __dadd_rn(__dsub_rn(two_t3, three_t2), 1.0), h10 = __dadd_rn(__dsub_rn(t3, two_t2), tt);
        const double h01 = __dsub_rn(three_t2, two_t3), h11 = __dsub_rn(t3, t2);
        const uint32_t i_hi = K - 3;  // inclusive; keys i and i+1 are regular for every i <= K-3
        auto term = [&](double v, double o) -> double {
            return TAME ? mape_term_tame(round_and_limit5_tame(v, vmin, vmax), o)
                        : mape_term(round_and_limit5_fast(v, vmin, vmax), o);
        };
        auto hermite = [&](double av, double ta, double bv, double tb) -> double {
            return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(av, h00), __dmul_rn(ta, h10)), __dmul_rn(bv, h01)),
                             __dmul_rn(tb, h11));
        };
        auto seg_err = [&](uint32_t i) -> double {
            const uint32_t xa = i * step;
            return term(hermite(d[xa], tang[i], d[xa + step], tang[i + 1]), d[xa + j]);
        };
        // NS adjacent segments per trip through running pointers: the NS - 1 inner keys and tangents are
        // loaded once for the two segments they belong to, and 3 * NS + 2 loads are in flight per thread.
        // The loop waits on memory, not on issue slots: with the same arithmetic, two segments per trip
        // took 1.00 ms per bench step, two with one-trip-ahead register prefetch 0.89, four 0.78 (B200).
        constexpr int NS = POLY_NS;
        const uint32_t nblk = i_hi / NS;  // segments 1 .. i_hi in blocks of NS
        const double *p = d + (size_t)(1 + NS * (b_lo + g)) * step, *tg = tang + 1 + NS * (b_lo + g);
        const double *po = p + j;  // this thread's samples
        const size_t stride = (size_t)NS * G * step;
        for (uint32_t q = b_lo + g; q < b_hi; q += G) {
            double kv[NS + 1], tv[NS + 1], o[NS], e[NS];
#pragma unroll
            for (int u = 0; u <= NS; u++) {
                kv[u] = p[(uint32_t)u * step];
                tv[u] = tg[u];
            }
#pragma unroll
            for (int u = 0; u < NS; u++) o[u] = po[(uint32_t)u * step];
#pragma unroll
            for (int u = 0; u < NS; u++) e[u] = term(hermite(kv[u], tv[u], kv[u + 1], tv[u + 1]), o[u]);
#pragma unroll
            for (int u = 0; u < NS; u++) acc += e[u];
            p += stride;
            po += stride;
            tg += NS * G;
        }
        if (tail)
            for (uint32_t i = nblk * NS + 1 + g; i <= i_hi; i += G) acc += seg_err(i);  // fewer than NS left over
    }
    return acc;
}

// MAPE (utils/error.rs:104-116) of one candidate step against the frame; block-wide.
// Catmull-Rom path: identical value arithmetic to poly_eval_at (same operations, same order).
// Thread t owns one offset j inside the segments (its Hermite basis values stay in registers) and
// walks over segments, so the inner loop has no table lookups, no index division and no branches;
// the per-key tangents come from a pre-pass.
// TAME selects the guard-free arithmetic of a tame frame (same values, fewer instructions).
template <bool TAME>
__device__ inline double poly_mape(const double *__restrict__ d, const PolyKeys &k, int ptype,
                                   double vmin, double vmax, const double *__restrict__ inv_d2,
                                   PolyWs ws, double *scratch) {
    const uint32_t N = k.N, step = k.step, K = k.K, T = blockDim.x, t = threadIdx.x;
    double acc = 0.0;
    auto pts = [&](uint32_t j) { return d[poly_pos(k, j)]; };
    if (ptype || step >= (uint32_t)POLY_MAXSTEP || K < 2) {
        for (uint32_t x = t; x < N; x += T) {
            double v = ptype ? idw_eval_at(k, x, pts, inv_d2) : poly_eval_at(k, x, pts);
            double o = d[x];
            acc += mape_term(round_and_limit5_fast(v, vmin, vmax), o);  // IDW values are not bounded like splines
        }
        double s = block_sum(acc, scratch);
        return __ddiv_rn(s, (double)N);
    }
    // ---- tangent at every interior key: (v[j+1] - v[j-1]) / (pos[j+1] - pos[j-1]) * step
    // (both segments that use it as a Catmull-Rom tangent are regular, i.e. `step` long)
    const double stepd = (double)step;
    double *__restrict__ tang = ws.slope;
    for (uint32_t j = 1 + t; j + 1 < K; j += T) {
        uint32_t pa = poly_pos(k, j - 1), pb = poly_pos(k, j + 1);
        tang[j] = __dmul_rn(__ddiv_rn(__dsub_rn(d[pb], d[pa]), __dsub_rn((double)pb, (double)pa)), stepd);
    }
    __syncthreads();
    if (K >= 4) acc += poly_cr_range<TAME>(d, k, vmin, vmax, tang, 0u, 0u, (K - 3) / POLY_NS, true);
    acc += poly_mape_ends(d, k, vmin, vmax);
    double s = block_sum(acc, scratch);
    return __ddiv_rn(s, (double)N);
}

// k_poly1 (queue-driven predecessor of k_poly1s, kept for A/B runs): the FIRST candidate step of a big bounded
// Catmull-Rom frame, cut into POLY_ITEM-sample work items so
// that the pass balances over the SMs (a wave of the bench fleet holds ~340 full frames for 296 CTA slots: whole
// frames leave a quarter of the SM time idle).  Item q of Q = poly_item_count(N) takes the NS-blocks
// [nblk * q / Q, nblk * (q + 1) / Q) of the step poly_frame tries first, the last item also the left-over segments
// and the Linear ends; its tangents live in shared memory.  The MAPE is a sum: the items' partial sums are added
// in item order by poly_frame, which goes on from there exactly as if it had evaluated the step itself.
template <bool TAME>
__device__ inline void poly_first_step_item(const double *__restrict__ d, uint32_t N, double vmin, double vmax, uint32_t q,
                                            double *part_out, double *tang_sm, double *scratch) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const PolyKeys k = poly_keys(N, poly_first_step(N));
    const uint32_t K = k.K, Q = poly_item_count(N), nblk = (K - 3) / POLY_NS;
    const uint32_t b_lo = (uint32_t)((uint64_t)nblk * q / Q), b_hi = (uint32_t)((uint64_t)nblk * (q + 1) / Q);
    const bool tail = q + 1 == Q;
    const uint32_t j_lo = 1 + POLY_NS * b_lo, j_hi = tail ? K - 2 : POLY_NS * b_hi + 1;  // keys whose tangent is used
    const double stepd = (double)k.step;
    for (uint32_t j = j_lo + t; j <= j_hi; j += T) {
        const uint32_t pa = poly_pos(k, j - 1), pb = poly_pos(k, j + 1);
        tang_sm[j - j_lo] = __dmul_rn(__ddiv_rn(__dsub_rn(d[pb], d[pa]), __dsub_rn((double)pb, (double)pa)), stepd);
    }
    __syncthreads();
    double acc = poly_cr_range<TAME>(d, k, vmin, vmax, tang_sm, j_lo, b_lo, b_hi, tail);
    if (tail) acc += poly_mape_ends(d, k, vmin, vmax);
    const double s = block_sum(acc, scratch);
    if (t == 0) *part_out = s;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// k_poly1s (default): the same first step without the per-item latency chain.
//
// ncu on the item kernel above (profiles/r2_p1_variants.md): 45 % of the warp-stall samples sit OUTSIDE the
// segment loop -- queue atomic -> item -> frame record -> keys -> tangents -> barrier, then block_sum -- a chain of
// four dependent global accesses per 32768-sample item.  Here nothing on an item's path is a dependent global access:
//   * k_plan, which decides per frame whether the Polynomial candidate runs at all, appends one self-contained
//     64-byte descriptor per work item to a compacted list (the items of constant frames never exist);
//   * k_poly1s walks that list with a STATIC schedule (CTA b takes entries b, b + grid, ...: the entries are all
//     ~32768 samples): descriptors arrive two items ahead and the next item's raw keys one item ahead by cp.async
//     (their tangents are computed from shared memory at the item boundary), the samples one trip ahead in
//     registers -- the first trip of the NEXT item included, loaded while the current item's sums are reduced --
//     and the result leaves as one partial sum per warp (no block-wide reduction).  Two barriers per item.
//   * the samples the four-segment blocks do not cover (left-over segments, Linear ends: ~300 per frame) are
//     evaluated by poly_frame when it adds up the frame's partial sums (poly_first_step_rest).
// Staging the samples through shared-memory cp.async rings instead (measured) removes every load stall and costs
// 26 more instructions per trip -- and instructions are what this loop is short of: a B200 FP64 instruction holds
// its scheduler's issue port for two cycles (tools/ubench/p1arith.cu), so a trip costs 2 * 76 + 63 cycles.
// step is 100 for every frame of >= 10000 samples: N / (N / 100) = 100 + floor((N mod 100) / (N / 100)).
struct P1Smem {
    double2 kt[POLY_ITEM_KEYS];          // (key value, tangent) of the current item's keys
    double raw[2][POLY_ITEM_KEYS + 2];   // raw keys j_lo - 1 .. j_hi + 1 of the current / the next item
    P1Item desc[4];                      // descriptors of the items k-1 .. k+2 of this CTA
};
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}
// this thread's NS samples of the trip that starts at p (one per segment of the block)
__device__ __forceinline__ void p1_load_trip(double (&o)[POLY_NS], const double *__restrict__ p) {
#pragma unroll
    for (int u = 0; u < POLY_NS; u++) o[u] = __ldg(p + (uint32_t)u * P1_STEP);
}
// One trip of k_poly1s' segment loop: the MAPE terms of this thread's NS samples o[] against the Hermite values of
// the NS segments whose (key, tangent) pairs start at kp.
// TAME frames (poly_tame) take the guard-free arithmetic, with one more saving: every sample has the sign of vmin,
// so the rounding addend copysign(pred(0.5), y) of round_half_away is the same for the whole frame (`half`).  A
// spline value that overshoots across zero gets the wrong-signed addend; its rounded value still lies on the far
// side of zero from [vmin, vmax] (|vmin|, |vmax| >= 1e-200) and is clamped to the same bound as the exact one.
template <bool TAME>
__device__ __forceinline__ double p1_trip(const double2 *kp, const double (&o)[POLY_NS], double vmin, double vmax, double half,
                                          double h00, double h10, double h01, double h11) {
    constexpr int NS = POLY_NS;
    double e[NS];
    double2 kv[NS + 1];
#pragma unroll
    for (int u = 0; u <= NS; u++) kv[u] = kp[u];
#pragma unroll
    for (int u = 0; u < NS; u++) {
        const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(kv[u].x, h00), __dmul_rn(kv[u].y, h10)), __dmul_rn(kv[u + 1].x, h01)),
                                   __dmul_rn(kv[u + 1].y, h11));
        if (TAME) {
            double out = div_1e5_int53(trunc(__dadd_rn(__dmul_rn(v, 100000.0), half)));
            out = out < vmin ? vmin : (out > vmax ? vmax : out);
            e[u] = mape_term_tame(out, o[u]);
        } else {
            e[u] = mape_term(round_and_limit5_fast(v, vmin, vmax), o[u]);
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < NS; u++) acc += e[u];
    return acc;
}
// The segment loop of one item for one thread (offset j of group g); returns its share of the MAPE sum.
// oA holds the samples of the item's first trip on entry (loaded while the previous item finished); inside the
// loop the next trip's samples are loaded one trip ahead, alternating between two register sets (two trips per
// round, no copies): the FP64 work of a trip covers the latency of the next trip's loads.
#ifndef P1_L2_AHEAD
#define P1_L2_AHEAD 3u
#endif
template <bool TAME>
__device__ __forceinline__ double p1_item_loop(const P1Item &I, const double2 *kp, double (&oA)[POLY_NS], bool active,
                                               uint32_t g, uint32_t j, double h00, double h10, double h01, double h11) {
    constexpr int NS = POLY_NS;
    constexpr size_t TRIP = (size_t)NS * P1_G * P1_STEP;  // samples between two trips of a thread
    const double vmin = I.vmin, vmax = I.vmax;
    const double half = copysign(0.49999999999999994, vmin);
    const uint32_t b_hi = I.b_hi;
    const double *po = I.d + (size_t)(1u + NS * (I.b_lo + g)) * P1_STEP + j;  // this thread's samples of the current trip
    double acc = 0.0;
    double oB[NS];
    // (the lanes beyond the fifth group make no trip but stay on the same path: barriers and shuffles are warp-wide)
    uint32_t qq = active ? I.b_lo + g : b_hi;
    for (; qq + P1_G < b_hi; qq += 2u * P1_G) {
        p1_load_trip(oB, po + TRIP);
        if (P1_L2_AHEAD && qq + (P1_L2_AHEAD + 1u) * P1_G < b_hi) {
            // the trips P1_L2_AHEAD and P1_L2_AHEAD + 1 from now: on their way into L2, no register held
#pragma unroll
            for (int u = 0; u < 2 * NS; u++)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(po + P1_L2_AHEAD * TRIP + (u >> 2) * TRIP + (uint32_t)(u & 3) * P1_STEP));
        }
        acc += p1_trip<TAME>(kp, oA, vmin, vmax, half, h00, h10, h01, h11);
        po += 2u * TRIP;
        if (qq + 2u * P1_G < b_hi) p1_load_trip(oA, po);
        acc += p1_trip<TAME>(kp + NS * P1_G, oB, vmin, vmax, half, h00, h10, h01, h11);
        kp += 2u * NS * P1_G;
    }
    if (qq < b_hi) acc += p1_trip<TAME>(kp, oA, vmin, vmax, half, h00, h10, h01, h11);  // odd trip count
    return acc;
}

// The part of the first step's MAPE sum that k_poly1s' four-segment blocks do not cover: the Catmull-Rom segments
// behind the last whole block (fewer than NS) through the per-sample arithmetic -- which yields the same values as
// the segment loop (poly_mape) -- and the Linear ends.  Returns this thread's share; all threads call.
__device__ inline double poly_first_step_rest(const double *__restrict__ d, const PolyKeys &k, double vmin, double vmax) {
    auto pts = [&](uint32_t jx) { return d[poly_pos(k, jx)]; };
    const uint32_t nblk = (k.K - 3) / POLY_NS, i0 = nblk * POLY_NS + 1, i_hi = k.K - 3;
    const uint32_t nleft = i_hi >= i0 ? (i_hi - i0 + 1) * k.step : 0u;
    double acc = 0.0;
    for (uint32_t e = threadIdx.x; e < nleft; e += blockDim.x) {
        const uint32_t x = i0 * k.step + e;
        acc += mape_term(round_and_limit5_fast(poly_eval_at(k, x, pts), vmin, vmax), d[x]);
    }
    return acc + poly_mape_ends(d, k, vmin, vmax);
}

// Polynomial::polynomial_to_data (polynomial.rs:342-373) + round_and_limit_f64 for a whole frame:
// out[x] for x < N from the K decoded key values pts[] -- the decompressor's hot loop.  Same value
// arithmetic as poly_eval_at (same operations, same order; the final quotient by 1e5 is the
// correctly rounded one, div_1e5), organised like poly_mape: thread t owns one offset inside the
// segments, tangents come from a pre-pass.  All threads call.
// smem / smem_cap: optional shared-memory scratch (doubles); when the keys and tangents of the frame
// fit (2 K <= smem_cap, e.g. every 131072-sample frame stored at step >= 25) they are served from
// there instead of the L2-resident global scratch.
__device__ inline void poly_expand(const double *__restrict__ pts_g, const PolyKeys &k, double vmin, double vmax,
                                   double *__restrict__ tang_g, double *__restrict__ out, double *smem = nullptr,
                                   uint32_t smem_cap = 0) {
    const uint32_t N = k.N, step = k.step, K = k.K, T = blockDim.x, t = threadIdx.x;
    const double *pts = pts_g;
    double *tang = tang_g;
    if (smem && 2u * K <= smem_cap) {
        for (uint32_t j = t; j < K; j += T) smem[j] = pts_g[j];
        pts = smem;
        tang = smem + K;
        __syncthreads();
    }
    auto pf = [&](uint32_t j) { return pts[j]; };
    // |stored range| <= 1e10: wherever the clamp lets the quotient through, |n| <= 1e15 < 2^53 and the one-step
    // form is the IEEE quotient (common.cuh: div_1e5_int53); beyond it the value is clamped away anyway
    const bool small = fmax(fabs(vmin), fabs(vmax)) <= 1e10;
    auto fin = [&](double v) {
        const double n = round_half_away(__dmul_rn(v, 100000.0));
        double o = small ? div_1e5_int53(n) : div_1e5(n);
        if (o < vmin) return vmin;
        if (o > vmax) return vmax;
        return o;
    };
    if (step >= (uint32_t)POLY_MAXSTEP || K < 4) {
        for (uint32_t x = t; x < N; x += T) out[x] = fin(poly_eval_at(k, x, pf));
        return;
    }
    const double stepd = (double)step;
    const uint32_t G = T / step, g = t / step, j = t - g * step;
    double h00 = 0.0, h10 = 0.0, h01 = 0.0, h11 = 0.0;
    if (g < G) {
        const double tt = __ddiv_rn((double)j, stepd);
        const double two_t = __dmul_rn(tt, 2.0), three_t = __dmul_rn(tt, 3.0);
        const double t2 = __dmul_rn(tt, tt), t3 = __dmul_rn(t2, tt);
        const double two_t3 = __dmul_rn(t2, two_t), two_t2 = __dmul_rn(tt, two_t), three_t2 = __dmul_rn(tt, three_t);
        h00 = __dadd_rn(__dsub_rn(two_t3, three_t2), 1.0);
        h10 = __dadd_rn(__dsub_rn(t3, two_t2), tt);
        h01 = __dsub_rn(three_t2, two_t3);
        h11 = __dsub_rn(t3, t2);
    }
    // tangent at key jj from its neighbours' values (both neighbours regular except for the appended last key)
    auto tangent = [&](uint32_t jj, double vm, double vp) {
        const uint32_t pa = poly_pos(k, jj - 1), pb = poly_pos(k, jj + 1);
        return __dmul_rn(__ddiv_rn(__dsub_rn(vp, vm), __dsub_rn((double)pb, (double)pa)), stepd);
    };
    if (smem && pts == pts_g && smem_cap >= 64u) {
        // Many keys (small steps): the keys do not fit shared memory at once.  The frame is expanded in
        // chunks of C segments whose keys and tangents are staged in shared memory, so that the inner loop
        // reads them at shared-memory latency instead of from the L2-resident scratch.
        const uint32_t C = (smem_cap - 8u) / 2u;  // segments per chunk
        double *sp = smem, *st = smem + C + 4u;   // sp[q] = pts[c0 - 1 + q], st[q] = tangent of key c0 + q
        for (uint32_t c0 = 1; c0 + 3 <= K; c0 += C) {
            const uint32_t ns = min(C, K - 2u - c0);  // Catmull-Rom segments c0 .. c0 + ns - 1
            __syncthreads();
            for (uint32_t q = t; q < ns + 3u; q += T) sp[q] = pts_g[c0 - 1u + q];  // keys c0-1 .. c0+ns+1
            __syncthreads();
            for (uint32_t q = t; q <= ns; q += T) st[q] = tangent(c0 + q, sp[q], sp[q + 2u]);  // keys c0 .. c0+ns (<= K-2)
            __syncthreads();
            if (g < G) {
#pragma unroll 4
                for (uint32_t q = g; q < ns; q += G) {
                    const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sp[q + 1u], h00), __dmul_rn(st[q], h10)),
                                                         __dmul_rn(sp[q + 2u], h01)),
                                               __dmul_rn(st[q + 1u], h11));
                    out[(c0 + q) * step + j] = fin(v);
                }
            }
        }
    } else {
        for (uint32_t jj = 1 + t; jj + 1 < K; jj += T) tang[jj] = tangent(jj, pts[jj - 1], pts[jj + 1]);
        __syncthreads();
        if (g < G) {
            // Catmull-Rom segments 1 .. K-3 (polynomial.rs:349): keys i and i+1 are regular
#pragma unroll 4
            for (uint32_t i = 1 + g; i + 3 <= K; i += G) {
                const double av = pts[i], bv = pts[i + 1];
                const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(av, h00), __dmul_rn(tang[i], h10)), __dmul_rn(bv, h01)),
                                           __dmul_rn(tang[i + 1], h11));
                out[i * step + j] = fin(v);
            }
        }
    }
    // the Linear ends: segment 0, segment K-2 (possibly irregular), the last sample
    const uint32_t start_last = (K - 2) * step, nA = step, nB = N - start_last;
    for (uint32_t e = t; e < nA + nB; e += T) {
        const uint32_t x = e < nA ? e : start_last + (e - nA);
        out[x] = fin(poly_eval_at(k, x, pf));
    }
}

// payload size of a Polynomial struct (polynomial.rs:54-87) with K points at `step`
__device__ inline uint32_t poly_payload_size(const double *__restrict__ d, const PolyKeys &k,
                                             int bitdepth, bool no_points, uint32_t *scratch) {
    if (no_points) return 1 + 1 + 1 + 16 + 1;
    uint32_t vb = 0;
    if (bitdepth == BD_U8)
        vb = k.K;
    else if (bitdepth == BD_F64)
        vb = 8 * k.K;
    else {
        uint32_t loc = 0;
        for (uint32_t j = threadIdx.x; j < k.K; j += blockDim.x)
            loc += value_bytes(d[poly_pos(k, j)], bitdepth);
        vb = block_sum_u32(loc, scratch);
    }
    return 1 + 1 + varint_len(k.K) + vb + 16 + 1;
}

// near-tie detector for `round3(e) < round4(cur)` (polynomial.rs:231,255)
__device__ inline bool poly_loop_near_tie(double cur, double target) {
    if (!(cur == cur) || isinf(cur)) return false;
    double c4 = cur * 10000.0;
    double fr = c4 - floor(c4);
    return fabs(fr - 0.5) < 1e-6 && fabs(round(c4) / 10000.0 - target) < 2.5e-4;
}

// Runs the reference's refinement loop for one frame. All threads of the CTA call.
// Writes poly_* fields of fw (thread 0).  `sh` = shared scratch (>= 40 doubles).
__device__ inline void poly_frame(const double *__restrict__ d, FrameWork *fw, double max_err,
                                  const double *__restrict__ inv_d2, double *sh, PolyWs ws,
                                  const double *__restrict__ first_parts, uint32_t parts_per_item) {
    const uint32_t N = fw->len;
    const double vmin = fw->vmin, vmax = fw->vmax;
    const int ptype = fw->poly_type;
    const int bitdepth = fw->bitdepth;
    uint32_t step = 1;
    double cur = 0.0;
    uint32_t it = 0;
    bool tie = false;
    bool no_points = (vmax == vmin);  // polynomial.rs:210,280 "Same max and min"
    const bool tame = poly_tame(vmin, vmax);
    const uint32_t baseline = (3 >= N / 100) ? 3 : N / 100;
    if (!no_points) {
        if (!fw->bounded) {
            // Polynomial::compress (polynomial.rs:307-314)
            step = N / baseline;
            if (step < 1) step = 1;
        } else {
            cur = max_err + 1.0;
            uint32_t jump = 0;
            const double target = round_f64_dec(max_err, 3);
            uint32_t prev_step = 0;
            double prev_err = 0.0;
            while (target < round_f64_dec(cur, 4)) {
                it++;
                uint32_t points = baseline + jump;
                step = N / points;
                if (step < 1) step = 1;
                PolyKeys k = poly_keys(N, step);
                if (step == prev_step) {
                    cur = prev_err;  // same keys -> same reconstruction -> same error
                } else if (step == 1 && it <= 22) {
                    cur = 0.0;  // value unused: the `len == data_len` exit below overrides it
                } else if (it == 1 && (fw->front_res & FRES_POLY1) && step == fw->poly_step) {
                    cur = fw->poly_err;  // k_front evaluated this step while it streamed the frame
                } else if (it == 1 && fw->poly_parts && ptype == 0) {
                    // k_poly1 / k_poly1s evaluated this step in poly_parts work items: add their partial sums in a
                    // fixed order (one per item, or -- k_poly1s -- one per warp of each item plus the samples its blocks leave out)
                    double sum = 0.0;
                    if (parts_per_item == 1u) {
                        for (uint32_t q = 0; q < fw->poly_parts; q++) sum += first_parts[fw->poly_part0 + q];
                    } else {
                        const uint32_t np = fw->poly_parts * parts_per_item;  // <= 4 * 16
                        double mine = poly_first_step_rest(d, k, vmin, vmax);
                        for (uint32_t e = threadIdx.x; e < np; e += blockDim.x) mine += first_parts[(size_t)fw->poly_part0 * parts_per_item + e];
                        sum = block_sum(mine, sh);
                    }
                    cur = __ddiv_rn(sum, (double)N);
                } else {
                    cur = tame ? poly_mape<true>(d, k, ptype, vmin, vmax, inv_d2, ws, sh)
                               : poly_mape<false>(d, k, ptype, vmin, vmax, inv_d2, ws, sh);
                }
                prev_step = step;
                prev_err = cur;
                tie = tie || poly_loop_near_tie(cur, target);
                if (it <= 17) {
                    uint32_t j = N / 10;
                    jump += j > 1 ? j : 1;
                } else if (it <= 22) {
                    uint32_t j = N / 100;
                    jump += j > 1 ? j : 1;
                } else if (target > round_f64_dec(cur, 4)) {
                    break;
                } else {
                    step = 1;  // compress_hinted(data, data_len): store everything
                    cur = 0.0;
                    break;
                }
                if (k.K == N) {
                    cur = 0.0;
                    break;
                }
            }
        }
    }
    PolyKeys k = poly_keys(N, step);
    uint32_t size = poly_payload_size(d, k, bitdepth, no_points, (uint32_t *)sh);
    if (threadIdx.x == 0) {
        fw->poly_step = step;
        fw->poly_npts = no_points ? 0 : k.K;
        fw->poly_size = size;
        fw->poly_iters = (uint16_t)it;
        fw->poly_tie = tie ? 1 : 0;
        fw->poly_err = cur;
        fw->poly_valid = 1;
    }
    __syncthreads();
}

// Emits the Polynomial payload (polynomial.rs:54-87) at `out`; all threads call.
// sh: >= 40 uint32 scratch.
__device__ inline void poly_emit(const double *__restrict__ d, const FrameWork *fw, uint8_t *out,
                                 uint32_t *sh) {
    const uint32_t N = fw->len;
    const int bitdepth = fw->bitdepth;
    const bool no_points = fw->poly_npts == 0;
    PolyKeys k = poly_keys(N, fw->poly_step);
    uint32_t K = no_points ? 0 : k.K;
    uint32_t hdr = 2 + varint_len(K);
    if (threadIdx.x == 0) {
        out[0] = fw->poly_type ? 1 : 0;  // PolynomialType variant (polynomial.rs:29-34)
        out[1] = (uint8_t)bitdepth;
        put_varint(out + 2, K);
    }
    uint32_t body = 0;
    if (bitdepth == BD_U8 || bitdepth == BD_F64) {
        uint32_t w = bitdepth == BD_U8 ? 1 : 8;
        for (uint32_t j = threadIdx.x; j < K; j += blockDim.x)
            put_value(out + hdr + (size_t)j * w, d[poly_pos(k, j)], bitdepth);
        body = K * w;
    } else {
        // varint-coded points: tile-wise exclusive scan of byte lengths
        uint32_t base = 0;
        for (uint32_t j0 = 0; j0 < K; j0 += blockDim.x) {
            uint32_t j = j0 + threadIdx.x;
            double v = j < K ? d[poly_pos(k, j)] : 0.0;
            uint32_t len = j < K ? value_bytes(v, bitdepth) : 0;
            uint32_t tot;
            uint32_t off = block_excl_scan_u32(len, sh, &tot);
            if (j < K) put_value(out + hdr + base + off, v, bitdepth);
            base += tot;
            __syncthreads();
        }
        body = base;
    }
    if (threadIdx.x == 0) {
        uint8_t *p = out + hdr + body;
        double mn = fw->vmin, mx = fw->vmax;
        put_bytes(p, &mn, 8);
        put_bytes(p + 8, &mx, 8);
        p[16] = (uint8_t)fw->poly_step;  // `step as u8`
    }
}

}  // namespace atsc
