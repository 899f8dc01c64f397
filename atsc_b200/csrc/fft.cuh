// fft.cuh -- batched f32 FFT engine + FFT compressor loop, device side.
//   FFT::compress_bounded  fft.rs:288-362     FFT::fft_trim        fft.rs:231-257
//   FFT::gibbs_sizing      fft.rs:184-204     get_mirrored_freqs   fft.rs:401-422
//   FFT::round             fft.rs:208-218     FFT::to_data         fft.rs:426-462
//
// Engine (prototype with the same index math: tools/fft_prototype.py):
//  * length L = 2^a 3^b.  Even L uses the real-input trick: complex length M = L/2.
//  * four-step M = M1 x M2 with both sub-lengths done in shared memory by a Stockham
//    DIF autosort, 32 transforms side by side (batch index = lane, so twiddles are warp
//    uniform and every shared access is unit stride).
//  * the spectrum stays in "permuted" order S[k1][k2] = Z[k1 + M1*k2]; the inverse consumes
//    that order directly, so there is no transpose pass.
//  * the inverse's first pass builds its input tile straight from the sparse top-k list.
#pragma once
#include "common.cuh"

namespace atsc {

constexpr int FFT_THREADS = 512;                                // CTA size of the FFT kernels: 2 CTAs per SM
constexpr int FFT_LANES = 16;                                   // transforms side by side in one tile
constexpr int FFT_SUB = 32 / FFT_LANES;                         // butterfly slots per warp
constexpr int FFT_SLOTS = (FFT_THREADS / 32) * FFT_SUB;         // butterfly slots per CTA (32: FftStage.dp/dq)
constexpr int FFT_TLEN = 320;                                   // max sub-FFT length
constexpr int FFT_FP = FFT_LANES + 1;                           // padded batch stride (float2 units)
constexpr int FFT_TILE_F2 = FFT_TLEN * FFT_FP;                  // float2 per tile buffer
constexpr int FFT_SMEM_BYTES = 2 * FFT_TILE_F2 * (int)sizeof(float2);  // 87,040 B -> two CTAs per SM
constexpr int FFT_SORT_CAP = 8192;                              // u64 entries the tile buffers can sort at once

// thread -> (butterfly slot, transform lane) of a tile
__device__ inline int fft_lane() { return threadIdx.x & (FFT_LANES - 1); }
__device__ inline int fft_slot() { return (threadIdx.x >> 5) * FFT_SUB + ((threadIdx.x & 31) / FFT_LANES); }
constexpr int FFT_KCAP = 16384;                                 // max list entries when compressing (pow2 for the sort)
constexpr int FFT_DEC_KCAP = 65536;                             // max entries accepted when decoding (pos is a u16)
constexpr int FFT_SCHED = 23;                                   // fft.rs:348-352
constexpr int TOPK_U = 16;                                      // key loads in flight per thread in the top-k passes

// one Stockham stage of a sub-FFT: everything the inner loop needs, computed on the host so
// the kernel does no integer division
struct FftStage {
    uint16_t r, m, s, tws;   // radix, n/r, stride, twiddle stride (len / n)
    uint16_t nbf, dp, dq, pad; // butterflies per transform (len / r); (nw / s, nw % s) for nw = 32 warps
    uint32_t magic;          // ceil(65536 / s): warp / s == (warp * magic) >> 16 for warp < 32, s <= 320
};

struct FftGeom {
    uint32_t L, real, M, M1, M2, Bn;
    uint32_t ns1, ns2;
    FftStage st1[12], st2[12];
    const float2 *twM;  // exp(-2 pi i j / M), j < M
    const float2 *tw1;  // exp(-2 pi i j / M1)
    const float2 *tw2;  // exp(-2 pi i j / M2)
    const float2 *twL;  // exp(-2 pi i j / L), j <= M
    const float2 *twL1; // exp(-2 pi i k1 / L), k1 < M1
    const float2 *twL2; // exp(-2 pi i M1 k2 / L), k2 < M2
    const float2 *twA;  // [M1][32]: exp(-2 pi i e f / M)
    const float2 *twB;  // [M2][32]
    const float2 *T4;   // [M2][M1] four-step twiddle exp(-2 pi i k1 n2 / M), column-major (fft2.cuh); nullptr = none
    const float2 *T4T;  // [M1][M2] the same table, row-major (inverse, fft2.cuh)
};

// radix RA of pass-1 stage 1 when the geometry has a probe (fft2.cuh: M = M1 x 243 with
// M1 = RA * RB in {288, 144, 72, 36}), else 0.  The probe's stage-1 sums are folds over RA
// contiguous chunks of RB * 243 complex elements, which k_front (front.cuh) accumulates.
__host__ __device__ inline uint32_t f2_fold_ra(uint32_t M1) {
    return M1 == 288u || M1 == 144u ? 16u : M1 == 72u ? 8u : M1 == 36u ? 4u : 0u;
}
// k_sfold (sfold.cuh): work items of SF_ITEM slots, SF_FOLD_SLOTS float4 (A, B) per frame in the wave's fold arena
constexpr uint32_t SF_ITEM = 1458;             // a third of a full frame's 18 * 243 slots
constexpr uint32_t SF_FOLD_SLOTS = 18u * 243u;
// slots of a frame's fold: RB * 243 with M1 = RA * RB (0: the geometry has no probe)
__host__ __device__ inline uint32_t sfold_slots(uint32_t M1) {
    const uint32_t RA = f2_fold_ra(M1);
    return RA ? (M1 / RA) * 243u : 0u;
}
__host__ __device__ inline uint32_t sfold_items(uint32_t M1) { return (sfold_slots(M1) + SF_ITEM - 1u) / SF_ITEM; }

// per-CTA-slot global workspace
struct FftWs {
    float2 *W;       // [Mmax]   four-step intermediate / permuted spectrum
    float2 *Xd;      // [Bnmax]  half spectrum, natural order
    uint32_t *keys;  // [Bnmax]  f32 bits of |X[k]|
    uint32_t *rank;  // [Bnmax]  list rank per bin (+1), 0 = not in list
    uint32_t *locD, *locM, *ovr;  // [FFT_DEC_KCAP]
    float2 *cD, *cM;              // [FFT_DEC_KCAP]
    FftEntry *dlist;              // [FFT_DEC_KCAP] decode-side entry list
};

__device__ inline float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ inline float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ inline float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ inline float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// ---------------------------------------------------------------------------------------
// small DFTs in registers (checked against np.fft in tools/radix_check.py)
// ---------------------------------------------------------------------------------------
template <bool INV>
__device__ inline float2 mul_i(float2 v) {  // v * (+i) for the inverse, v * (-i) for the forward transform
    return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
}
template <bool INV>
__device__ inline void dft3(float2 &a0, float2 &a1, float2 &a2) {
    const float c = 0.86602540378443864676f;
    float2 t = cadd(a1, a2), d = csub(a1, a2);
    float2 m = make_float2(a0.x - 0.5f * t.x, a0.y - 0.5f * t.y);
    float2 rot = mul_i<INV>(make_float2(c * d.x, c * d.y));
    a0 = cadd(a0, t);
    a1 = cadd(m, rot);
    a2 = csub(m, rot);
}
template <bool INV>
__device__ inline void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}
template <bool INV>
__device__ inline void dft8(float2 *a) {  // in place, natural order out
    const float r = 0.70710678118654752440f;
    float2 l0 = cadd(a[0], a[4]), l1 = cadd(a[1], a[5]), l2 = cadd(a[2], a[6]), l3 = cadd(a[3], a[7]);
    float2 h0 = csub(a[0], a[4]), h1 = csub(a[1], a[5]), h2 = csub(a[2], a[6]), h3 = csub(a[3], a[7]);
    // h_j *= W8^j : W8 = exp(-+ i pi/4)
    float2 i1 = mul_i<INV>(h1), i3 = mul_i<INV>(h3);
    h1 = make_float2(r * (h1.x + i1.x), r * (h1.y + i1.y));
    h2 = mul_i<INV>(h2);
    h3 = make_float2(r * (i3.x - h3.x), r * (i3.y - h3.y));
    dft4<INV>(l0, l1, l2, l3);
    dft4<INV>(h0, h1, h2, h3);
    a[0] = l0;
    a[1] = h0;
    a[2] = l1;
    a[3] = h1;
    a[4] = l2;
    a[5] = h2;
    a[6] = l3;
    a[7] = h3;
}
template <bool INV>
__device__ inline void dft9(float2 *a) {  // in place, natural order out
    // W9^k = exp(-+ 2 pi i k / 9)
    const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;
    const float c2 = 0.17364817766693034885f, s2 = 0.98480775301220805937f;
    const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;
    const float2 w1 = make_float2(c1, INV ? s1 : -s1), w2 = make_float2(c2, INV ? s2 : -s2),
                 w4 = make_float2(c4, INV ? s4 : -s4);
    dft3<INV>(a[0], a[3], a[6]);  // t[0][k1] in a[0], a[3], a[6]
    dft3<INV>(a[1], a[4], a[7]);  // t[1][k1]
    dft3<INV>(a[2], a[5], a[8]);  // t[2][k1]
    a[4] = cmul(a[4], w1);
    a[7] = cmul(a[7], w2);
    a[5] = cmul(a[5], w2);
    a[8] = cmul(a[8], w4);
    dft3<INV>(a[0], a[1], a[2]);  // k1 = 0 -> y[0], y[3], y[6]
    dft3<INV>(a[3], a[4], a[5]);  // k1 = 1 -> y[1], y[4], y[7]
    dft3<INV>(a[6], a[7], a[8]);  // k1 = 2 -> y[2], y[5], y[8]
    float2 y1 = a[3], y2 = a[6], y3 = a[1], y5 = a[7], y6 = a[2], y7 = a[5];
    a[1] = y1;
    a[2] = y2;
    a[3] = y3;
    a[5] = y5;
    a[6] = y6;
    a[7] = y7;
}

// One Stockham DIF stage of radix R for the 32 transforms of a tile.
//   y[q + s*(R*p + u)] = DFT_R(x[q + s*(p + t*m)], t < R)[u] * W_n^{p*u}
// LAST: the stage's own twiddles are all 1 (m == 1).  FUSE (only with LAST): multiply the result
// element e by the four-step twiddle W_M^{+-(e*(base+lane))} = twM[e*base] * twEF[e][lane].
template <int R, bool INV, bool LAST, bool FUSE>
__device__ inline void fft_stage(const float2 *x, float2 *y, const FftStage S, const float2 *__restrict__ tw,
                                 int nb, uint32_t base, const float2 *__restrict__ twM,
                                 const float2 *__restrict__ twEF) {
    const int warp = fft_slot(), lane = fft_lane(), nw = FFT_SLOTS;
    const int m = S.m, s = S.s, tws = S.tws, nbf = S.nbf;
    if (lane >= nb) return;
    int p = (warp * (int)S.magic) >> 16, q = warp - p * s;
    const int dp = S.dp, dq = S.dq;
    const int xs = s * m * FFT_FP, ys = s * FFT_FP;
    for (int b = warp; b < nbf; b += nw, p += dp, q += dq) {
        if (q >= s) {
            q -= s;
            p++;
        }
        const float2 *xi = x + (q + s * p) * FFT_FP + lane;
        float2 a[R];
#pragma unroll
        for (int t = 0; t < R; t++) a[t] = xi[t * xs];
        if (R == 2) {
            float2 t0 = cadd(a[0], a[1]);
            a[1] = csub(a[0], a[1]);
            a[0] = t0;
        } else if (R == 3) {
            dft3<INV>(a[0], a[1], a[2]);
        } else if (R == 4) {
            dft4<INV>(a[0], a[1], a[2], a[3]);
        } else if (R == 8) {
            dft8<INV>(a);
        } else {
            dft9<INV>(a);
        }
        const int eo = q + s * R * p;  // output element of leg 0
        float2 *yo = y + eo * FFT_FP + lane;
        if (!LAST) {
#pragma unroll
            for (int u = 1; u < R; u++) {
                float2 w = tw[u * p * tws];
                a[u] = INV ? cmulc(a[u], w) : cmul(a[u], w);
            }
        }
        if (FUSE) {
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int e = eo + u * s;
                float2 w = cmul(twM[(uint32_t)e * base], twEF[e * 32 + lane]);
                a[u] = INV ? cmulc(a[u], w) : cmul(a[u], w);
            }
        }
#pragma unroll
        for (int u = 0; u < R; u++) yo[u * ys] = a[u];
    }
}

template <bool INV, bool LAST, bool FUSE>
__device__ inline void fft_stage_any(const float2 *x, float2 *y, const FftStage S, const float2 *__restrict__ tw,
                                     int nb, uint32_t base, const float2 *__restrict__ twM,
                                     const float2 *__restrict__ twEF) {
    switch (S.r) {
        case 2: fft_stage<2, INV, LAST, FUSE>(x, y, S, tw, nb, base, twM, twEF); break;
        case 3: fft_stage<3, INV, LAST, FUSE>(x, y, S, tw, nb, base, twM, twEF); break;
        case 4: fft_stage<4, INV, LAST, FUSE>(x, y, S, tw, nb, base, twM, twEF); break;
        case 8: fft_stage<8, INV, LAST, FUSE>(x, y, S, tw, nb, base, twM, twEF); break;
        default: fft_stage<9, INV, LAST, FUSE>(x, y, S, tw, nb, base, twM, twEF); break;
    }
}

// ---------------------------------------------------------------------------------------
// 32-wide batched Stockham DIF FFT of length `len` in shared memory.
// Element e of transform f lives at x[e*FFT_FP + f].  Returns the buffer with the result.
// twM != nullptr: the four-step twiddle (see fft_stage) is fused into the last stage.
// ---------------------------------------------------------------------------------------
template <bool INV>
__device__ inline float2 *tile_fft(float2 *x, float2 *y, int len, const FftStage *stg, int ns,
                                   const float2 *__restrict__ tw, int nb, uint32_t base = 0,
                                   const float2 *__restrict__ twM = nullptr,
                                   const float2 *__restrict__ twEF = nullptr) {
    (void)len;
    for (int st = 0; st < ns; st++) {
        const FftStage S = stg[st];
        if (st + 1 < ns)
            fft_stage_any<INV, false, false>(x, y, S, tw, nb, base, twM, twEF);
        else if (twM)
            fft_stage_any<INV, true, true>(x, y, S, tw, nb, base, twM, twEF);
        else
            fft_stage_any<INV, true, false>(x, y, S, tw, nb, base, twM, twEF);
        __syncthreads();
        float2 *t = x;
        x = y;
        y = t;
    }
    return x;
}

// rows r0..r0+nb-1 of the [M1][M2] global matrix <-> tile (element = column index)
__device__ inline void tile_load_rows(float2 *x, const float2 *__restrict__ W, int M2, int r0, int nb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nb) {
        const float2 *row = W + (size_t)(r0 + warp) * M2;
        for (int e = lane; e < M2; e += 32) x[e * FFT_FP + warp] = row[e];
    }
    __syncthreads();
}
__device__ inline void tile_store_rows(const float2 *x, float2 *__restrict__ W, int M2, int r0, int nb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nb) {
        float2 *row = W + (size_t)(r0 + warp) * M2;
        for (int e = lane; e < M2; e += 32) row[e] = x[e * FFT_FP + warp];
    }
    __syncthreads();
}
// columns c0..c0+nb-1 (element = row index)
__device__ inline void tile_load_cols(float2 *x, const float2 *__restrict__ W, int M1, int M2, int c0, int nb) {
    const int warp = fft_slot(), lane = fft_lane(), nw = FFT_SLOTS;
    if (lane < nb)
        for (int e = warp; e < M1; e += nw) x[e * FFT_FP + lane] = W[(size_t)e * M2 + c0 + lane];
    __syncthreads();
}
__device__ inline void tile_store_cols(const float2 *x, float2 *__restrict__ W, int M1, int M2, int c0, int nb) {
    const int warp = fft_slot(), lane = fft_lane(), nw = FFT_SLOTS;
    if (lane < nb)
        for (int e = warp; e < M1; e += nw) W[(size_t)e * M2 + c0 + lane] = x[e * FFT_FP + lane];
    __syncthreads();
}

// padded ("gibbs sized", fft.rs:184-204) sample j of the frame, as the f32 the reference feeds
// to the FFT (fft.rs:221-228)
__device__ inline double padded_sample(const double *__restrict__ d, uint32_t N, uint32_t prefix, uint32_t j) {
    uint32_t i = j < prefix ? 0u : j - prefix;
    if (i >= N) i = N - 1;
    return d[i];
}

// L2 prefetch of the samples a column tile (columns c0..c0+31, all M1 rows) will read; issued one
// tile ahead so the DRAM latency overlaps the current tile's butterflies
__device__ inline void prefetch_col_tile(const double *__restrict__ d, uint32_t N, uint32_t prefix, int M1, int M2,
                                         int c0, int real) {
    if (c0 >= M2) return;
    const int per_row = (FFT_LANES * (real ? 2 : 1) * 8 + 127) / 128;  // 128-byte lines per row of the tile
    for (int i = threadIdx.x; i < M1 * per_row; i += blockDim.x) {
        int e = i / per_row, ln = i - e * per_row;
        uint32_t n = (uint32_t)e * M2 + c0;
        uint32_t j = (real ? 2 * n : n) + 16 * ln;
        uint32_t idx = j < prefix ? 0u : j - prefix;
        if (idx < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + idx));
    }
}

// ---------------------------------------------------------------------------------------
// forward transform of the padded frame -> Xd[k], keys[k], k < Bn
// ---------------------------------------------------------------------------------------
__device__ inline void fft_forward(const double *__restrict__ d, uint32_t N, uint32_t prefix,
                                   const FftGeom &g, FftWs ws, float2 *sm) {
    float2 *bufA = sm, *bufB = sm + FFT_TILE_F2;
    const int M1 = g.M1, M2 = g.M2;
    const int warp = fft_slot(), lane = fft_lane(), nw = FFT_SLOTS;
    // pass 1: column tiles; input built from the samples
    for (int c0 = 0; c0 < M2; c0 += FFT_LANES) {
        int nb = min(FFT_LANES, M2 - c0);
        prefetch_col_tile(d, N, prefix, M1, M2, c0 + FFT_LANES, g.real);
        if (lane < nb) {
            for (int e = warp; e < M1; e += nw) {
                uint32_t n = (uint32_t)e * M2 + c0 + lane;
                float2 z;
                if (g.real) {
                    z.x = (float)padded_sample(d, N, prefix, 2 * n);
                    z.y = (float)padded_sample(d, N, prefix, 2 * n + 1);
                } else {
                    z.x = (float)padded_sample(d, N, prefix, n);
                    z.y = 0.f;
                }
                bufA[e * FFT_FP + lane] = z;
            }
        }
        __syncthreads();
        float2 *res = tile_fft<false>(bufA, bufB, M1, g.st1, g.ns1, g.tw1, nb, (uint32_t)c0, g.twM, g.twA);
        tile_store_cols(res, ws.W, M1, M2, c0, nb);
    }
    // pass 2: row tiles, in place
    for (int r0 = 0; r0 < M1; r0 += FFT_LANES) {
        int nb = min(FFT_LANES, M1 - r0);
        tile_load_rows(bufA, ws.W, M2, r0, nb);
        float2 *res = tile_fft<false>(bufA, bufB, M2, g.st2, g.ns2, g.tw2, nb);
        tile_store_rows(res, ws.W, M2, r0, nb);
    }
    __threadfence_block();
    __syncthreads();
    // post-process: half spectrum X[k], k = 0..L/2, and its |X[k]| keys.
    // Real-input trick: entries are produced in *storage* order i = k1*M2 + k2 (k = k1 + M1*k2) so
    // both Z[k] and its partner Z[M-k] are read with unit stride; entry M holds the Nyquist bin.
    // Odd L: natural order (only the two small odd lengths 243 / 2187 take this path).
    const uint32_t M = g.M;
    if (g.real) {
        for (uint32_t i = threadIdx.x; i <= M; i += blockDim.x) {
            float2 X;
            if (i == M) {
                float2 Z0 = ws.W[0];
                X = make_float2(Z0.x - Z0.y, 0.f);  // X[M] = Re Z0 - Im Z0
            } else {
                uint32_t k1 = i / (uint32_t)M2, k2 = i - k1 * (uint32_t)M2;
                float2 Zk = ws.W[i];
                // partner M-k: (M1-k1, M2-1-k2) for k1 > 0, (0, M2-k2) for k1 == 0 < k2, itself for k == 0
                uint32_t pi = k1 ? (M1 - k1) * (uint32_t)M2 + ((uint32_t)M2 - 1 - k2) : (k2 ? (uint32_t)M2 - k2 : 0u);
                float2 Zm = ws.W[pi];
                Zm.y = -Zm.y;  // conj
                float2 sum = cadd(Zk, Zm), dif = csub(Zk, Zm);
                float2 w = cmul(g.twL1[k1], g.twL2[k2]);  // exp(-2 pi i (k1 + M1 k2) / L)
                float2 tt = cmul(w, dif);
                // X = 0.5 * (sum - i * tt)
                X.x = 0.5f * (sum.x + tt.y);
                X.y = 0.5f * (sum.y - tt.x);
            }
            ws.Xd[i] = X;
            // Complex<f32>::norm() == hypotf; via f64 sqrt to stay correctly rounded
            double nr = sqrt((double)X.x * (double)X.x + (double)X.y * (double)X.y);
            ws.keys[i] = __float_as_uint((float)nr);
        }
    } else {
        for (uint32_t k = threadIdx.x; k < g.Bn; k += blockDim.x) {
            float2 X = ws.W[(size_t)(k % M1) * M2 + k / M1];
            ws.Xd[k] = X;
            double nr = sqrt((double)X.x * (double)X.x + (double)X.y * (double)X.y);
            ws.keys[k] = __float_as_uint((float)nr);
        }
    }
    __syncthreads();
}

// array index of ws.Xd / ws.keys -> spectrum bin (see fft_forward)
__device__ inline uint32_t fft_bin_of(uint32_t i, uint32_t pM, uint32_t pM1, uint32_t pM2) {
    if (pM == 0 || i >= pM) return i;
    uint32_t k1 = i / pM2;
    return k1 + pM1 * (i - k1 * pM2);
}

// ---------------------------------------------------------------------------------------
// radix-select helper: bins are scanned from the largest digit down; returns the digit d with
//   #(keys with a larger digit) < remaining <= that + hist[d]        (uniform across the CTA)
// ---------------------------------------------------------------------------------------
__device__ inline void find_digit(const uint32_t *hist, uint32_t nbins, uint32_t remaining, uint32_t *sh,
                                  uint32_t *d, uint32_t *above, uint32_t *cnt) {
    const uint32_t t = threadIdx.x, per = nbins >= blockDim.x ? nbins / blockDim.x : 1;  // <= 8
    uint32_t loc[8], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < 8; j++) {
        uint32_t pos = t * per + j;  // position in descending-digit order
        loc[j] = (j < per && pos < nbins) ? hist[nbins - 1 - pos] : 0u;
        sum += loc[j];
    }
    uint32_t tot;
    uint32_t excl = block_excl_scan_u32(sum, sh, &tot);
    if (excl < remaining && remaining <= excl + sum) {
        uint32_t acc = excl;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) {
            if (j < per && acc < remaining && remaining <= acc + loc[j]) {
                sh[102] = nbins - 1 - (t * per + j);
                sh[103] = acc;
                sh[106] = loc[j];
            }
            acc += loc[j];
        }
    }
    __syncthreads();
    *d = sh[102];
    *above = sh[103];
    *cnt = sh[106];
    __syncthreads();
}

// sort key of a spectrum bin: |z| (f32 bits) descending, then bin ascending; unique per bin
__device__ inline unsigned long long fft_composite(uint32_t key, uint32_t bin) {
    return ((unsigned long long)key << 20) | (unsigned long long)(0xFFFFFu - bin);
}

// number of bins with a non-zero |z| (fft_trim stops at the first exact zero, fft.rs:249-252)
__device__ inline uint32_t fft_count_nonzero(uint32_t Bn, FftWs ws, uint32_t *sh) {
    uint32_t zeros = 0;
    constexpr int U = 16;  // loads in flight per thread: the passes over the keys are latency bound
    for (uint32_t b0 = threadIdx.x; b0 < Bn; b0 += U * blockDim.x) {
        uint32_t k[U];
#pragma unroll
        for (int u = 0; u < U; u++) k[u] = (b0 + u * blockDim.x < Bn) ? ws.keys[b0 + u * blockDim.x] : 1u;
#pragma unroll
        for (int u = 0; u < U; u++) zeros += k[u] == 0u;
    }
    return Bn - block_sum_u32(zeros, sh);
}

// ---------------------------------------------------------------------------------------
// top-k of the half spectrum by |z| (fft.rs:231-257): writes the entries of rank
// [done, done + want) in descending (|z|, then lower bin) order to list[done ..].
//   Cprev  : composite of entry done-1 (~0 for the first chunk); *Clast <- composite of the last one
//   (pM, pM1, pM2): storage order of ws.keys (pM = 0: natural order)
// A radix select over the 52-bit composite (12-bit digits) narrows the bucket that holds rank
// done+want until "everything at or above that bucket" fits the shared-memory sort; those
// candidates are compacted (unordered), bitonic-sorted, and the first `want` kept.
// Returns true if equal |z| values sit on both sides of the final cut (BinaryHeap pops equal
// norms in unspecified order -> near-tie).
// ---------------------------------------------------------------------------------------
__device__ inline bool fft_topk_chunk(uint32_t Bn, FftWs ws, uint32_t done, uint32_t want, unsigned long long Cprev,
                                      FftEntry *list, unsigned long long *sm64, uint32_t *sh, uint32_t pM,
                                      uint32_t pM1, uint32_t pM2, unsigned long long *Clast) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    uint32_t *hist = (uint32_t *)sm64;  // <= 4096 bins (phase-local reuse of the tile buffers)
    const uint32_t target = done + want;
    unsigned long long prefix = 0;
    uint32_t remaining = target;
    int shift = 52;
    // ---- radix select on the composite
    const int shifts[5] = {40, 28, 16, 4, 0};
    for (int lv = 0; lv < 5; lv++) {
        const int bits = lv == 4 ? 4 : 12;
        const uint32_t nbins = 1u << bits;
        const int hi_shift = shift;  // bits above this are already fixed to `prefix`
        shift = shifts[lv];
        for (uint32_t i = t; i < nbins; i += T) hist[i] = 0;
        __syncthreads();
        for (uint32_t b0 = t; b0 < Bn; b0 += TOPK_U * T) {
            uint32_t k[TOPK_U];
#pragma unroll
            for (int u = 0; u < TOPK_U; u++) k[u] = (b0 + u * T < Bn) ? ws.keys[b0 + u * T] : 0u;
#pragma unroll
            for (int u = 0; u < TOPK_U; u++)
                if (k[u] != 0u) {
                    uint32_t bin = shift < 20 ? fft_bin_of(b0 + u * T, pM, pM1, pM2) : 0u;
                    unsigned long long c = fft_composite(k[u], bin);
                    if (hi_shift >= 52 || (c >> hi_shift) == prefix) atomicAdd(&hist[(uint32_t)(c >> shift) & (nbins - 1)], 1u);
                }
        }
        __syncthreads();
        uint32_t dsel, above, cnt;
        find_digit(hist, nbins, remaining, sh, &dsel, &above, &cnt);
        remaining -= above;
        prefix = (prefix << bits) | dsel;
        // candidates if the whole bucket is taken = (#composites above the bucket) + cnt - done
        uint32_t ncand = (target - remaining) + cnt - done;
        uint32_t P = 1;
        while (P < ncand) P <<= 1;
        if (P <= (uint32_t)FFT_SORT_CAP) break;
    }
    const unsigned long long Tlow = prefix << shift;  // smallest composite of the boundary bucket
    // ---- compaction (unordered; the sort orders it)
    unsigned long long *S = sm64;
    __syncthreads();
    if (t == 0) sh[107] = 0;
    __syncthreads();
    for (uint32_t b0 = t; b0 < Bn; b0 += TOPK_U * T) {
        uint32_t k[TOPK_U];
#pragma unroll
        for (int u = 0; u < TOPK_U; u++) k[u] = (b0 + u * T < Bn) ? ws.keys[b0 + u * T] : 0u;
#pragma unroll
        for (int u = 0; u < TOPK_U; u++)
            if (k[u] != 0u) {
                unsigned long long c = fft_composite(k[u], fft_bin_of(b0 + u * T, pM, pM1, pM2));
                if (c >= Tlow && c < Cprev) {
                    uint32_t pos = atomicAdd(&sh[107], 1u);
                    if (pos < (uint32_t)FFT_SORT_CAP) S[pos] = c;
                }
            }
    }
    __syncthreads();
    const uint32_t nsel = min(sh[107], (uint32_t)FFT_SORT_CAP);
    uint32_t P = 1;
    while (P < nsel) P <<= 1;
    for (uint32_t i = nsel + t; i < P; i += T) S[i] = 0ull;
    __syncthreads();
    // ---- bitonic sort, descending
    for (uint32_t k2 = 2; k2 <= P; k2 <<= 1) {
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = t; i < P; i += T) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = S[i], b = S[ixj];
                    bool desc = (i & k2) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        S[i] = b;
                        S[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    const uint32_t take = min(want, nsel);
    for (uint32_t r = t; r < take; r += T) {
        uint32_t bin = 0xFFFFFu - (uint32_t)(S[r] & 0xFFFFFull);
        uint32_t i = bin;  // array index of this bin (inverse of fft_bin_of)
        if (pM && bin < pM) i = (bin % pM1) * pM2 + bin / pM1;
        float2 X = ws.Xd[i];
        FftEntry e;
        e.bin = bin;
        e.re = X.x;
        e.im = X.y;
        list[done + r] = e;
    }
    *Clast = take ? S[take - 1] : Cprev;
    // equal |z| on both sides of the cut?  (the whole boundary bucket was taken, so the next
    // candidate is visible whenever one with the same top bits exists)
    bool tie = take > 0 && nsel > take && (S[take - 1] >> 20) == (S[take] >> 20);
    if (take > 0 && nsel == take) {
        // nothing beyond the cut among the candidates: look for an equal key below the bucket
        const uint32_t kcut = (uint32_t)(S[take - 1] >> 20);
        uint32_t eq = 0;
        for (uint32_t b = t; b < Bn; b += T) eq += ws.keys[b] == kcut;
        uint32_t eq_all = block_sum_u32(eq, sh);
        uint32_t eq_sel = 0;
        for (uint32_t r = t; r < take; r += T) eq_sel += (uint32_t)(S[r] >> 20) == kcut;
        // entries of earlier chunks with the same key are above the cut as well
        for (uint32_t r = t; r < done; r += T) {
            uint32_t bin = list[r].bin, i = bin;
            if (pM && bin < pM) i = (bin % pM1) * pM2 + bin / pM1;
            eq_sel += ws.keys[i] == kcut;
        }
        uint32_t eq_in = block_sum_u32(eq_sel, sh);
        tie = eq_all > eq_in;
    }
    __syncthreads();
    return tie;
}

// builds list[0..K) for K = min(kmax, nonzero bins) in chunks the shared-memory sort can hold;
// cut_tie[i] (i < ncuts) tells whether a cut after cuts[i] entries separates equal |z|.
// Returns K.  After the call sm64 holds the LAST chunk only.
__device__ inline uint32_t fft_topk(uint32_t Bn, FftWs ws, uint32_t kmax, FftEntry *list,
                                    unsigned long long *sm64, uint32_t *sh, bool *tie_at_cut,
                                    uint32_t pM, uint32_t pM1, uint32_t pM2) {
    const uint32_t nz = fft_count_nonzero(Bn, ws, sh);
    const uint32_t K = min(min(kmax, nz), (uint32_t)FFT_KCAP);
    *tie_at_cut = false;
    if (K == 0) return 0;
    constexpr uint32_t CHUNK = FFT_SORT_CAP / 2 + FFT_SORT_CAP / 4;  // leaves room for the over-selected bucket
    uint32_t done = 0;
    unsigned long long Cprev = ~0ull;
    while (done < K) {
        uint32_t want = min(K - done, CHUNK);
        unsigned long long Clast;
        bool tie = fft_topk_chunk(Bn, ws, done, want, Cprev, list, sm64, sh, pM, pM1, pM2, &Clast);
        done += want;
        Cprev = Clast;
        if (done >= K) *tie_at_cut = tie;
    }
    return K;
}

// equal |z| on both sides of a cut after the first c list entries (c < K)?
__device__ inline bool fft_cut_splits_tie_list(const FftEntry *list, uint32_t c, uint32_t K) {
    if (c == 0 || c >= K) return false;
    FftEntry a = list[c - 1], b = list[c];
    double na = sqrt((double)a.re * a.re + (double)a.im * a.im), nb = sqrt((double)b.re * b.re + (double)b.im * b.im);
    return (float)na == (float)nb;
}

// ---------------------------------------------------------------------------------------
// per-entry scatter coefficients for the inverse (see tools/fft_prototype.py)
// ---------------------------------------------------------------------------------------
// set_ovr=false: ws.ovr was filled by the caller (decode path: "last entry per position wins")
__device__ inline void fft_prepare_entries(const FftGeom &g, FftWs ws, const FftEntry *list,
                                           uint32_t K, bool alias, bool set_ovr = true) {
    const uint32_t M = g.M, M1 = g.M1, L = g.L;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    if (alias) {
        for (uint32_t b = t; b < g.Bn; b += T) ws.rank[b] = 0;
        __syncthreads();
        for (uint32_t r = t; r < K; r += T) ws.rank[list[r].bin] = r + 1;
        __syncthreads();
    }
    for (uint32_t r = t; r < K; r += T) {
        FftEntry e = list[r];
        uint32_t p = alias ? (e.bin & 0xFFFFu) : e.bin;  // `pos as u16` (fft.rs:242)
        uint32_t ov = 0xFFFFFFFFu;
        if (alias) {
            uint32_t partner = e.bin ^ 0x10000u;
            if (partner < g.Bn) {
                uint32_t pr = ws.rank[partner];
                if (pr) ov = pr - 1;
            }
        }
        if (set_ovr) ws.ovr[r] = ov;
        float2 z = make_float2(e.re, e.im);
        uint32_t lD = 0xFFFFFFFFu, lM = 0xFFFFFFFFu;
        float2 cD = make_float2(0.f, 0.f), cM = make_float2(0.f, 0.f);
        if (g.real) {
            if (p == 0) {
                lD = 0;
                cD = make_float2(z.x, z.x);  // re * (1 + i)
            } else if (p == M) {
                lM = 0;
                cM = make_float2(z.x, -z.x);  // re * (1 - i) at k = 0
            } else if (p < M) {
                float2 w = g.twL[p];  // (cos, -sin)
                float c = w.x, s = -w.y;
                cD = cmul(z, make_float2(1.f - s, c));
                lD = ((p % M1) << 16) | (p / M1);
                uint32_t k = M - p;
                cM = cmul(make_float2(z.x, -z.y), make_float2(1.f + s, c));
                lM = ((k % M1) << 16) | (k / M1);
            }
        } else {
            if (p == 0) {
                lD = 0;
                cD = make_float2(z.x, 0.f);
            } else if (p < L) {
                lD = ((p % M1) << 16) | (p / M1);
                cD = z;
                uint32_t k = L - p;
                lM = ((k % M1) << 16) | (k / M1);
                cM = make_float2(z.x, -z.y);
            }
        }
        ws.locD[r] = lD;
        ws.locM[r] = lM;
        ws.cD[r] = cD;
        ws.cM[r] = cM;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// inverse transform of the first `c` list entries; Epi(j, value) is called once for every
// time index j < L with the unnormalised real output (all threads participate).
// ---------------------------------------------------------------------------------------
// (pf_d, pf_N, pf_prefix): samples the epilogue will read, prefetched one tile ahead (nullptr: none)
template <class Epi>
__device__ inline void fft_inverse(const FftGeom &g, FftWs ws, uint32_t c, float2 *sm, Epi epi,
                                   const double *pf_d = nullptr, uint32_t pf_N = 0, uint32_t pf_prefix = 0) {
    float2 *bufA = sm, *bufB = sm + FFT_TILE_F2;
    const int M1 = g.M1, M2 = g.M2;
    const int warp = fft_slot(), lane = fft_lane(), nw = FFT_SLOTS;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    // pass 1: row tiles built from the sparse list
    for (int r0 = 0; r0 < M1; r0 += FFT_LANES) {
        int nb = min(FFT_LANES, M1 - r0);
        for (uint32_t i = t; i < (uint32_t)M2 * FFT_FP; i += T) bufA[i] = make_float2(0.f, 0.f);
        __syncthreads();
        for (uint32_t r = t; r < c; r += T) {
            uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;  // overwritten by a later aliased entry (fft.rs:411-420)
            uint32_t l = ws.locD[r];
            if (l != 0xFFFFFFFFu) {
                uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < (uint32_t)FFT_LANES) bufA[(l & 0xFFFFu) * FFT_FP + f] = ws.cD[r];
            }
        }
        __syncthreads();
        for (uint32_t r = t; r < c; r += T) {
            uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;
            uint32_t l = ws.locM[r];
            if (l != 0xFFFFFFFFu) {
                uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < (uint32_t)FFT_LANES) {
                    float2 *q = &bufA[(l & 0xFFFFu) * FFT_FP + f];
                    *q = cadd(*q, ws.cM[r]);
                }
            }
        }
        __syncthreads();
        float2 *res = tile_fft<true>(bufA, bufB, M2, g.st2, g.ns2, g.tw2, nb, (uint32_t)r0, g.twM, g.twB);
        tile_store_rows(res, ws.W, M2, r0, nb);
    }
    __threadfence_block();
    __syncthreads();
    // pass 2: column tiles + epilogue
    if (pf_d) prefetch_col_tile(pf_d, pf_N, pf_prefix, M1, M2, 0, g.real);
    for (int c0 = 0; c0 < M2; c0 += FFT_LANES) {
        int nb = min(FFT_LANES, M2 - c0);
        if (pf_d) prefetch_col_tile(pf_d, pf_N, pf_prefix, M1, M2, c0 + FFT_LANES, g.real);
        tile_load_cols(bufA, ws.W, M1, M2, c0, nb);
        float2 *res = tile_fft<true>(bufA, bufB, M1, g.st1, g.ns1, g.tw1, nb);
        if (lane < nb) {
            for (int e = warp; e < M1; e += nw) {
                uint32_t n = (uint32_t)e * M2 + c0 + lane;
                float2 v = res[e * FFT_FP + lane];
                if (g.real) {
                    epi(2 * n, v.x);
                    epi(2 * n + 1, v.y);
                } else {
                    epi(n, v.x);
                }
            }
        }
        __syncthreads();
    }
}

// FFT::round (fft.rs:208-218) of `re / len_f32`
__device__ inline double fft_round(float x, float vminf, float vmaxf) {
    double out = round5_exact((double)x);
    if (out > (double)vmaxf) return (double)vmaxf;
    if (out < (double)vminf) return (double)vminf;
    return out;
}

// same value through the FMA-refined quotient (poly.cuh div_1e5), for the refinement loop
__device__ inline double fft_round_fast(float x, float vminf, float vmaxf) {
    double n = round_half_away(__dmul_rn((double)x, 100000.0));
    double out = div_1e5(n);
    if (out > (double)vmaxf) return (double)vmaxf;
    if (out < (double)vminf) return (double)vminf;
    return out;
}

// payload size of an FFT struct with the first c entries (fft.rs:119-130)
__host__ __device__ inline uint32_t fft_payload_size(uint32_t c, uint32_t n_small_pos) {
    return 1 + varint_len(c) + 11 * c - 2 * n_small_pos + 8;
}

// ---------------------------------------------------------------------------------------
// direct O(n^2) path for frames shorter than 128 samples (no padding, fft.rs:305-309)
// ---------------------------------------------------------------------------------------
__device__ inline float2 unit_root(uint32_t num, uint32_t den, bool inverse) {
    double s, c;
    sincospi(2.0 * (double)(num % den) / (double)den, &s, &c);
    return make_float2((float)c, (float)(inverse ? s : -s));
}

}  // namespace atsc
