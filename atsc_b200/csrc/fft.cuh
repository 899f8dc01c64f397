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

constexpr int FFT_TLEN = 320;                                   // max sub-FFT length
constexpr int FFT_FP = 33;                                      // padded batch stride (float2 units)
constexpr int FFT_TILE_F2 = FFT_TLEN * FFT_FP;                  // float2 per tile buffer
constexpr int FFT_SMEM_BYTES = 2 * FFT_TILE_F2 * (int)sizeof(float2);  // 168,960 B
constexpr int FFT_KCAP = 16384;                                 // max list entries when compressing (pow2 for the sort)
constexpr int FFT_DEC_KCAP = 65536;                             // max entries accepted when decoding (pos is a u16)
constexpr int FFT_SCHED = 23;                                   // fft.rs:348-352

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
};

// per-CTA-slot global workspace
struct FftWs {
    float2 *W;       // [Mmax]   four-step intermediate / permuted spectrum
    float2 *Xd;      // [Bnmax]  half spectrum, natural order
    uint32_t *keys;  // [Bnmax]  f32 bits of |X[k]|
    uint32_t *rank;  // [Bnmax]  list rank per bin (+1), 0 = not in list
    uint32_t *locD, *locM, *ovr;  // [FFT_DEC_KCAP]
    float2 *cD, *cM;              // [FFT_DEC_KCAP]
    FftEntry *dlist;              // [FFT_DEC_KCAP] decode-side entry list
    double *w;                    // [MAX_FRAME + 8] 1 / sample for the MAPE terms of the refinement loop
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
// 32-wide batched Stockham DIF FFT of length `len` in shared memory.
// Element e of transform f lives at x[e*FFT_FP + f].  Returns the buffer with the result.
// ---------------------------------------------------------------------------------------
template <bool INV>
__device__ inline float2 *tile_fft(float2 *x, float2 *y, int len, const FftStage *stg, int ns,
                                   const float2 *__restrict__ tw, int nb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    (void)len;
    for (int st = 0; st < ns; st++) {
        const FftStage S = stg[st];
        const int r = S.r, m = S.m, s = S.s, tws = S.tws, nbf = S.nbf;
        if (lane < nb) {
            // butterfly b = p * s + q; (p, q) advance incrementally (nw = 32 warps)
            int p = (warp * (int)S.magic) >> 16, q = warp - p * s;
            const int dp = S.dp, dq = S.dq;
            for (int b = warp; b < nbf; b += nw, p += dp, q += dq) {
                if (q >= s) {
                    q -= s;
                    p++;
                }
                const float2 *xi = x + (q + s * p) * FFT_FP + lane;
                float2 *yo = y + (q + s * r * p) * FFT_FP + lane;
                const int xs = s * m * FFT_FP;  // input stride between radix legs
                const int ys = s * FFT_FP;      // output stride between radix legs
                if (r == 4) {
                    float2 a0 = xi[0], a1 = xi[xs], a2 = xi[2 * xs], a3 = xi[3 * xs];
                    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
                    // forward: -i*t3 ; inverse: +i*t3
                    float2 jt3 = INV ? make_float2(-t3.y, t3.x) : make_float2(t3.y, -t3.x);
                    float2 b0 = cadd(t0, t2), b1 = cadd(t1, jt3), b2 = csub(t0, t2), b3 = csub(t1, jt3);
                    float2 w1 = tw[p * tws], w2 = tw[2 * p * tws], w3 = tw[3 * p * tws];
                    yo[0] = b0;
                    yo[ys] = INV ? cmulc(b1, w1) : cmul(b1, w1);
                    yo[2 * ys] = INV ? cmulc(b2, w2) : cmul(b2, w2);
                    yo[3 * ys] = INV ? cmulc(b3, w3) : cmul(b3, w3);
                } else if (r == 3) {
                    float2 a0 = xi[0], a1 = xi[xs], a2 = xi[2 * xs];
                    float2 t = cadd(a1, a2), d = csub(a1, a2);
                    float2 mm = make_float2(a0.x - 0.5f * t.x, a0.y - 0.5f * t.y);
                    const float c = 0.86602540378443864676f;
                    float2 rot = INV ? make_float2(-c * d.y, c * d.x) : make_float2(c * d.y, -c * d.x);
                    float2 b0 = cadd(a0, t), b1 = cadd(mm, rot), b2 = csub(mm, rot);
                    float2 w1 = tw[p * tws], w2 = tw[2 * p * tws];
                    yo[0] = b0;
                    yo[ys] = INV ? cmulc(b1, w1) : cmul(b1, w1);
                    yo[2 * ys] = INV ? cmulc(b2, w2) : cmul(b2, w2);
                } else {  // r == 2
                    float2 a0 = xi[0], a1 = xi[xs];
                    float2 w1 = tw[p * tws];
                    float2 b1 = csub(a0, a1);
                    yo[0] = cadd(a0, a1);
                    yo[ys] = INV ? cmulc(b1, w1) : cmul(b1, w1);
                }
            }
        }
        __syncthreads();
        float2 *t = x;
        x = y;
        y = t;
    }
    return x;
}

// multiply tile element (e, f) by W_M^{+-(e * (base + f))} = twM[e*base] * twEF[e][f]
template <bool INV>
__device__ inline void tile_twiddle(float2 *x, int len, int nb, uint32_t base,
                                    const float2 *__restrict__ twM,
                                    const float2 *__restrict__ twEF) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane < nb) {
        for (int e = warp; e < len; e += nw) {
            float2 w = cmul(twM[(uint32_t)e * base], twEF[e * 32 + lane]);
            float2 v = x[e * FFT_FP + lane];
            x[e * FFT_FP + lane] = INV ? cmulc(v, w) : cmul(v, w);
        }
    }
    __syncthreads();
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane < nb)
        for (int e = warp; e < M1; e += nw) x[e * FFT_FP + lane] = W[(size_t)e * M2 + c0 + lane];
    __syncthreads();
}
__device__ inline void tile_store_cols(const float2 *x, float2 *__restrict__ W, int M1, int M2, int c0, int nb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
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
    const int per_row = real ? 4 : 2;  // 128-byte lines per row: 32 columns x (2 or 1) doubles
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // pass 1: column tiles; input built from the samples
    for (int c0 = 0; c0 < M2; c0 += 32) {
        int nb = min(32, M2 - c0);
        prefetch_col_tile(d, N, prefix, M1, M2, c0 + 32, g.real);
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
        float2 *res = tile_fft<false>(bufA, bufB, M1, g.st1, g.ns1, g.tw1, nb);
        tile_twiddle<false>(res, M1, nb, (uint32_t)c0, g.twM, g.twA);
        tile_store_cols(res, ws.W, M1, M2, c0, nb);
    }
    // pass 2: row tiles, in place
    for (int r0 = 0; r0 < M1; r0 += 32) {
        int nb = min(32, M1 - r0);
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
    const uint32_t t = threadIdx.x, per = nbins >= blockDim.x ? nbins / blockDim.x : 1;
    uint32_t loc[4], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        uint32_t pos = t * per + j;  // position in descending-digit order
        loc[j] = (j < per && pos < nbins) ? hist[nbins - 1 - pos] : 0u;
        sum += loc[j];
    }
    uint32_t tot;
    uint32_t excl = block_excl_scan_u32(sum, sh, &tot);
    if (excl < remaining && remaining <= excl + sum) {
        uint32_t acc = excl;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
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

// ---------------------------------------------------------------------------------------
// top-K of the half spectrum by |z| (fft.rs:231-257), descending, ties by lower bin.
// Writes list[0..K) and returns K = min(kmax, #nonzero bins).  (pM, pM1, pM2) describe the
// storage order of ws.keys (pM = 0: natural order).  sm64 must hold FFT_KCAP u64.
//
// Two 12-bit histogram passes fix the top 24 bits of the K-th largest key; every key at or
// above that 24-bit bucket is compacted (a few more than K), sorted, and the first K kept.
// Only when the boundary bucket is so crowded that the candidates would not fit does a third
// pass resolve the low 8 bits exactly.
// ---------------------------------------------------------------------------------------
__device__ inline uint32_t fft_topk(uint32_t Bn, FftWs ws, uint32_t kmax, FftEntry *list,
                                    unsigned long long *sm64, uint32_t *sh, bool *tie_at_cut,
                                    uint32_t pM, uint32_t pM1, uint32_t pM2) {
    const uint32_t T = blockDim.x, t = threadIdx.x;
    uint32_t *hist = (uint32_t *)sm64;  // 4096 bins (phase-local reuse of the big smem region)
    *tie_at_cut = false;
    // ---- level 1: top 12 bits (+ count of zero bins)
    for (uint32_t i = t; i < 4096; i += T) hist[i] = 0;
    __syncthreads();
    uint32_t zeros = 0;
    for (uint32_t b0 = t; b0 < Bn; b0 += 4 * T) {
        uint32_t k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = (b0 + u * T < Bn) ? ws.keys[b0 + u * T] : 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (k[u] != 0xFFFFFFFFu) {
                zeros += k[u] == 0u;
                atomicAdd(&hist[k[u] >> 20], 1u);
            }
    }
    const uint32_t nz = Bn - block_sum_u32(zeros, sh);
    const uint32_t K = min(min(kmax, nz), (uint32_t)FFT_KCAP);
    if (K == 0) return 0;
    uint32_t d1, above1, cnt1;
    find_digit(hist, 4096, K, sh, &d1, &above1, &cnt1);
    uint32_t remaining = K - above1;
    // ---- level 2: next 12 bits inside bucket d1
    for (uint32_t i = t; i < 4096; i += T) hist[i] = 0;
    __syncthreads();
    for (uint32_t b0 = t; b0 < Bn; b0 += 4 * T) {
        uint32_t k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = (b0 + u * T < Bn) ? ws.keys[b0 + u * T] : 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (k[u] != 0xFFFFFFFFu && (k[u] >> 20) == d1) atomicAdd(&hist[(k[u] >> 8) & 0xFFFu], 1u);
    }
    __syncthreads();
    uint32_t d2, above2, cnt2;
    find_digit(hist, 4096, remaining, sh, &d2, &above2, &cnt2);
    remaining -= above2;  // how many of the cnt2 keys in the boundary bucket belong to the top K
    const uint32_t T24 = (d1 << 12) | d2;
    const uint32_t Kover = K - remaining + cnt2;  // candidates if the whole bucket is taken
    uint32_t Pover = 1;
    while (Pover < Kover) Pover <<= 1;
    unsigned long long *S = sm64;
    uint32_t nsel = K, P;
    bool tie = false;
    __syncthreads();
    if (Pover <= (uint32_t)FFT_KCAP) {
        // ---- over-select: all keys whose top 24 bits are >= T24 (unordered; the sort orders them)
        if (t == 0) sh[107] = 0;
        __syncthreads();
        for (uint32_t b0 = t; b0 < Bn; b0 += 4 * T) {
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = (b0 + u * T < Bn) ? ws.keys[b0 + u * T] : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k[u] != 0u && (k[u] >> 8) >= T24) {
                    uint32_t pos = atomicAdd(&sh[107], 1u);
                    uint32_t bin = fft_bin_of(b0 + u * T, pM, pM1, pM2);
                    if (pos < (uint32_t)FFT_KCAP)
                        S[pos] = ((unsigned long long)k[u] << 32) | ((unsigned long long)(0xFFFFFu - bin) << 12);
                }
        }
        __syncthreads();
        nsel = min(sh[107], (uint32_t)FFT_KCAP);
        P = 1;
        while (P < nsel) P <<= 1;
    } else {
        // ---- crowded boundary bucket: resolve the low 8 bits, then take exactly K
        for (uint32_t i = t; i < 256; i += T) hist[i] = 0;
        __syncthreads();
        for (uint32_t b = t; b < Bn; b += T) {
            uint32_t key = ws.keys[b];
            if ((key >> 8) == T24) atomicAdd(&hist[key & 255u], 1u);
        }
        __syncthreads();
        uint32_t d3, above3, eq_total;
        find_digit(hist, 256, remaining, sh, &d3, &above3, &eq_total);
        const uint32_t take_eq = remaining - above3;
        const uint32_t Tkey = (T24 << 8) | d3;
        tie = eq_total > take_eq;
        P = 1;
        while (P < K) P <<= 1;
        // equal |z| may straddle the cut: take the lowest array indices first (ordered compaction)
        uint32_t base = 0, eqbase = 0;
        for (uint32_t b0 = 0; b0 < Bn; b0 += T) {
            uint32_t b = b0 + t;
            uint32_t key = b < Bn ? ws.keys[b] : 0u;
            bool gt = b < Bn && key > Tkey;
            bool eq = b < Bn && key == Tkey;
            uint32_t eqtot;
            uint32_t eqrank = eqbase + block_excl_scan_u32(eq ? 1u : 0u, sh, &eqtot);
            __syncthreads();
            bool sel = gt || (eq && eqrank < take_eq);
            uint32_t tot;
            uint32_t pos = base + block_excl_scan_u32(sel ? 1u : 0u, sh, &tot);
            if (sel) {
                uint32_t bin = fft_bin_of(b, pM, pM1, pM2);
                S[pos] = ((unsigned long long)key << 32) | ((unsigned long long)(0xFFFFFu - bin) << 12);
            }
            base += tot;
            eqbase += eqtot;
            __syncthreads();
        }
        nsel = K;
    }
    for (uint32_t i = nsel + t; i < P; i += T) S[i] = 0ull;
    __syncthreads();
    // bitonic sort, descending: (|z| desc, bin asc)
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
    // equal |z| on both sides of the K cut?
    if (nsel > K && (uint32_t)(S[K - 1] >> 32) == (uint32_t)(S[K] >> 32)) tie = true;
    for (uint32_t r = t; r < K; r += T) {
        uint32_t bin = 0xFFFFFu - (uint32_t)((S[r] >> 12) & 0xFFFFFull);
        // array index of this bin (inverse of fft_bin_of)
        uint32_t i = bin;
        if (pM && bin < pM) i = (bin % pM1) * pM2 + bin / pM1;
        float2 X = ws.Xd[i];
        FftEntry e;
        e.bin = bin;
        e.re = X.x;
        e.im = X.y;
        list[r] = e;
    }
    *tie_at_cut = tie;
    __syncthreads();
    return K;
}

// after fft_topk: S (sm64) still holds the sorted composite keys; true if a cut after the
// first c entries separates two bins with equal |z|
__device__ inline bool fft_cut_splits_tie(const unsigned long long *S, uint32_t c, uint32_t K) {
    return c > 0 && c < K && (uint32_t)(S[c - 1] >> 32) == (uint32_t)(S[c] >> 32);
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    // pass 1: row tiles built from the sparse list
    for (int r0 = 0; r0 < M1; r0 += 32) {
        int nb = min(32, M1 - r0);
        for (uint32_t i = t; i < (uint32_t)M2 * FFT_FP; i += T) bufA[i] = make_float2(0.f, 0.f);
        __syncthreads();
        for (uint32_t r = t; r < c; r += T) {
            uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;  // overwritten by a later aliased entry (fft.rs:411-420)
            uint32_t l = ws.locD[r];
            if (l != 0xFFFFFFFFu) {
                uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < 32u) bufA[(l & 0xFFFFu) * FFT_FP + f] = ws.cD[r];
            }
        }
        __syncthreads();
        for (uint32_t r = t; r < c; r += T) {
            uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;
            uint32_t l = ws.locM[r];
            if (l != 0xFFFFFFFFu) {
                uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < 32u) {
                    float2 *q = &bufA[(l & 0xFFFFu) * FFT_FP + f];
                    *q = cadd(*q, ws.cM[r]);
                }
            }
        }
        __syncthreads();
        float2 *res = tile_fft<true>(bufA, bufB, M2, g.st2, g.ns2, g.tw2, nb);
        tile_twiddle<true>(res, M2, nb, (uint32_t)r0, g.twM, g.twB);
        tile_store_rows(res, ws.W, M2, r0, nb);
    }
    __threadfence_block();
    __syncthreads();
    // pass 2: column tiles + epilogue
    if (pf_d) prefetch_col_tile(pf_d, pf_N, pf_prefix, M1, M2, 0, g.real);
    for (int c0 = 0; c0 < M2; c0 += 32) {
        int nb = min(32, M2 - c0);
        if (pf_d) prefetch_col_tile(pf_d, pf_N, pf_prefix, M1, M2, c0 + 32, g.real);
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

// same, for the refinement loop only: `* 1e-5` instead of the true `/ 1e5` (<= 1 ulp off; the error
// it feeds is compared against thresholds)
__device__ inline double fft_round_fast(float x, double o, float vminf, float vmaxf) {
    double n = round(__dmul_rn((double)x, 100000.0));
    double out = n * 1e-5;
    if (fabs(out - o) <= fabs(o) * 4.5e-16) out = __ddiv_rn(n, 100000.0);  // exact zero terms stay exact
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
