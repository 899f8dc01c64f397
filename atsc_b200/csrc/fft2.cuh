// fft2.cuh -- register-radix forward FFT engine for frames whose padded length is
// L = 2^a * 3^7 (a = 2..6: every power-of-two frame of 8192..131072 samples, fft.rs:184-204).
//
//   fft_forward (fft.cuh) restated for speed; same outputs, same storage order:
//     Xd[i], keys[i] with i = k1*M2 + k2 for bin k = k1 + M1*k2 < M, entry M = Nyquist bin.
//
// Real-input trick (complex length M = L/2 = M1 x 243), four-step, but every sub-FFT is only TWO
// Stockham stages whose butterflies (radix 16/18, 27/9) live entirely in registers:
//   pass 1  columns n2: samples -> registers (radix RA) -> smem -> registers (radix RB) -> W[n2][k1]
//           (four-step twiddle applied on the way out; W is column-major so pass 2 reads it coalesced)
//   pass 2  rows k1:    W -> registers (radix 27) -> smem -> registers (radix 9).  Row k1 is
//           processed together with row M1-k1 by the same thread, which therefore holds Z[k] and
//           Z[M-k] and finishes the real-input post-process without another exchange.
// Index math validated in tools/fft2_prototype.py.
#pragma once
#include "fft.cuh"

namespace atsc {

constexpr int F2_THREADS = 288;   // 9 warps; two CTAs per SM (<= 112 registers per thread)
constexpr int F2_TC = 32;         // columns per pass-1 tile
constexpr int F2_PAIRS = 16;      // row pairs per pass-2 tile (32 rows)
constexpr int F2_M2 = 243;
constexpr int F2_M1MAX = 288;
constexpr int F2_SMEM_F2 = F2_TC * (F2_M1MAX + 1);  // >= 32 * 243
constexpr int F2_SMEM_BYTES = F2_SMEM_F2 * (int)sizeof(float2);  // 73,984 B

// exp(-+ 2 pi i j / R) as immediates (j is a compile-time constant after unrolling)
template <int R, bool INV>
__device__ __forceinline__ float2 root_c(int j) {
    if (R == 4) {
        constexpr float C[4] = {1.0f, 0.0f, -1.0f, 0.0f};
        constexpr float S[4] = {0.0f, 1.0f, 0.0f, -1.0f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    if (R == 8) {
        constexpr float C[8] = {1.0f, 0.707106781186548f, 0.0f, -0.707106781186547f, -1.0f, -0.707106781186548f, 0.0f, 0.707106781186547f};
        constexpr float S[8] = {0.0f, 0.707106781186547f, 1.0f, 0.707106781186548f, 0.0f, -0.707106781186547f, -1.0f, -0.707106781186548f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    if (R == 9) {
        constexpr float C[9] = {1.0f, 0.766044443118978f, 0.17364817766693f, -0.5f, -0.939692620785908f, -0.939692620785908f, -0.5f, 0.17364817766693f, 0.766044443118978f};
        constexpr float S[9] = {0.0f, 0.642787609686539f, 0.984807753012208f, 0.866025403784439f, 0.342020143325669f, -0.342020143325669f, -0.866025403784438f, -0.984807753012208f, -0.64278760968654f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    if (R == 16) {
        constexpr float C[16] = {1.0f, 0.923879532511287f, 0.707106781186548f, 0.38268343236509f, 0.0f, -0.38268343236509f, -0.707106781186547f, -0.923879532511287f, -1.0f, -0.923879532511287f, -0.707106781186548f, -0.38268343236509f, 0.0f, 0.38268343236509f, 0.707106781186547f, 0.923879532511287f};
        constexpr float S[16] = {0.0f, 0.38268343236509f, 0.707106781186547f, 0.923879532511287f, 1.0f, 0.923879532511287f, 0.707106781186548f, 0.38268343236509f, 0.0f, -0.38268343236509f, -0.707106781186547f, -0.923879532511287f, -1.0f, -0.923879532511287f, -0.707106781186548f, -0.38268343236509f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    if (R == 18) {
        constexpr float C[18] = {1.0f, 0.939692620785908f, 0.766044443118978f, 0.5f, 0.17364817766693f, -0.17364817766693f, -0.5f, -0.766044443118978f, -0.939692620785908f, -1.0f, -0.939692620785908f, -0.766044443118978f, -0.5f, -0.17364817766693f, 0.17364817766693f, 0.499999999999999f, 0.766044443118978f, 0.939692620785908f};
        constexpr float S[18] = {0.0f, 0.342020143325669f, 0.642787609686539f, 0.866025403784439f, 0.984807753012208f, 0.984807753012208f, 0.866025403784439f, 0.642787609686539f, 0.342020143325669f, 0.0f, -0.342020143325669f, -0.642787609686539f, -0.866025403784438f, -0.984807753012208f, -0.984807753012208f, -0.866025403784439f, -0.64278760968654f, -0.342020143325669f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    if (R == 27) {
        constexpr float C[27] = {1.0f, 0.973044870579824f, 0.893632640323412f, 0.766044443118978f, 0.597158591702786f, 0.396079766039157f, 0.17364817766693f, -0.058144828910476f, -0.28680323271109f, -0.5f, -0.686241637868733f, -0.835487811412936f, -0.939692620785908f, -0.993238357741943f, -0.993238357741943f, -0.939692620785909f, -0.835487811412936f, -0.686241637868734f, -0.5f, -0.286803232711091f, -0.058144828910476f, 0.173648177666931f, 0.396079766039157f, 0.597158591702786f, 0.766044443118978f, 0.893632640323412f, 0.973044870579824f};
        constexpr float S[27] = {0.0f, 0.23061587074244f, 0.448799180200462f, 0.642787609686539f, 0.802123192755044f, 0.918216106880274f, 0.984807753012208f, 0.998308158271268f, 0.957989512315489f, 0.866025403784439f, 0.727373641573049f, 0.549508978070806f, 0.342020143325669f, 0.11609291412523f, -0.11609291412523f, -0.342020143325668f, -0.549508978070806f, -0.727373641573049f, -0.866025403784438f, -0.957989512315489f, -0.998308158271268f, -0.984807753012208f, -0.918216106880274f, -0.802123192755044f, -0.64278760968654f, -0.448799180200462f, -0.23061587074244f};
        return make_float2(C[j], INV ? S[j] : -S[j]);
    }
    return make_float2(1.f, 0.f);
}

// a *= exp(-+ 2 pi i j / R); trivial rotations cost no multiplies
template <int R, bool INV>
__device__ __forceinline__ float2 rot_c(float2 a, int j) {
    j %= R;
    if (j == 0) return a;
    if (2 * j == R) return make_float2(-a.x, -a.y);
    if (4 * j == R) return mul_i<INV>(a);
    if (4 * j == 3 * R) return mul_i<!INV>(a);
    return cmul(a, root_c<R, INV>(j));
}

// Natural-order in-place DFT of a[0], a[S], ..., a[(R-1)*S] (all indices compile-time constants).
template <int R, int S, bool INV>
struct DftS;

template <int S, bool INV>
struct DftS<1, S, INV> {
    static __device__ __forceinline__ void run(float2 *) {}
};
template <int S, bool INV>
struct DftS<2, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) {
        float2 t = cadd(a[0], a[S]);
        a[S] = csub(a[0], a[S]);
        a[0] = t;
    }
};
template <int S, bool INV>
struct DftS<3, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft3<INV>(a[0], a[S], a[2 * S]); }
};
template <int S, bool INV>
struct DftS<4, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft4<INV>(a[0], a[S], a[2 * S], a[3 * S]); }
};

// Cooley-Tukey R = R1 * R2 in registers: input n = n1*R2 + n2, output k = k1 + R1*k2
template <int R1, int R2, int S, bool INV>
__device__ __forceinline__ void dft_comp(float2 *a) {
    constexpr int R = R1 * R2;
#pragma unroll
    for (int n2 = 0; n2 < R2; n2++) DftS<R1, S * R2, INV>::run(a + n2 * S);  // -> t[n2][k1] at (k1*R2 + n2)
#pragma unroll
    for (int k1 = 1; k1 < R1; k1++)
#pragma unroll
        for (int n2 = 1; n2 < R2; n2++) a[(k1 * R2 + n2) * S] = rot_c<R, INV>(a[(k1 * R2 + n2) * S], n2 * k1);
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) DftS<R2, S, INV>::run(a + k1 * R2 * S);  // -> X[k1 + R1*k2] at (k1*R2 + k2)
    float2 t[R];
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++)
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) t[k1 + R1 * k2] = a[(k1 * R2 + k2) * S];
#pragma unroll
    for (int k = 0; k < R; k++) a[k * S] = t[k];
}
template <int S, bool INV>
struct DftS<8, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft_comp<2, 4, S, INV>(a); }
};
template <int S, bool INV>
struct DftS<9, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft_comp<3, 3, S, INV>(a); }
};
template <int S, bool INV>
struct DftS<16, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft_comp<4, 4, S, INV>(a); }
};
template <int S, bool INV>
struct DftS<18, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft_comp<2, 9, S, INV>(a); }
};
template <int S, bool INV>
struct DftS<27, S, INV> {
    static __device__ __forceinline__ void run(float2 *a) { dft_comp<3, 9, S, INV>(a); }
};

// ---------------------------------------------------------------------------------------
// The two stage-1 outputs the PROBE needs (q = 1 and q = RA-1 of the radix-RA butterfly over
// a[t] = z[t * RB*243 + m] = r_t + i q_t) are DEFINED through two sequential folds over t = 0 .. RA-1,
//     A = sum_t r_t w^t,   B = sum_t q_t w^t,   w = exp(-2 pi i / RA)      (one fold_acc per term),
//     s_1 = A + i B,       s_{RA-1} = conj(A) + i conj(B)                   (fold_out).
// The chunks t are CONTIGUOUS pieces of the padded frame, so k_front (front.cuh) accumulates A and
// B while the frame streams through shared memory once; the full transform and the stand-alone
// probe below evaluate the same expressions, so the bins of the probed rows are bit-identical
// wherever they are computed.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fold_acc(float4 &ab, float r, float q, float2 w) {
    ab.x = fmaf(r, w.x, ab.x);
    ab.y = fmaf(r, w.y, ab.y);
    ab.z = fmaf(q, w.x, ab.z);
    ab.w = fmaf(q, w.y, ab.w);
}
__device__ __forceinline__ void fold_out(float4 ab, float2 &s1, float2 &sR) {
    s1 = make_float2(__fsub_rn(ab.x, ab.w), __fadd_rn(ab.y, ab.z));
    sR = make_float2(__fadd_rn(ab.x, ab.w), __fsub_rn(ab.z, ab.y));
}
template <int RA>
__device__ __forceinline__ void fold_pair(const float2 *a, float2 &s1, float2 &sR) {
    float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < RA; t++) fold_acc(ab, a[t].x, a[t].y, root_c<RA, false>(t));
    fold_out(ab, s1, sR);
}

// padded ("gibbs sized") complex element n of the frame: (x[2n], x[2n+1]) as f32 (fft.rs:184-228).
// `vec`: d + (2n - prefix) is 16-byte aligned for every n, so interior elements take one 128-bit load.
__device__ __forceinline__ float2 f2_load_z(const double *__restrict__ d, int N, int prefix, int n, bool vec) {
    int i0 = 2 * n - prefix, i1 = i0 + 1;
    if (vec && i0 >= 0 && i1 < N) {
        const double2 v = __ldcs(reinterpret_cast<const double2 *>(d + i0));  // last use of the frame in this wave
        return make_float2((float)v.x, (float)v.y);
    }
    i0 = min(max(i0, 0), N - 1);
    i1 = min(max(i1, 0), N - 1);
    return make_float2((float)__ldg(d + i0), (float)__ldg(d + i1));
}

// ---------------------------------------------------------------------------------------
// pass 1: M1 = RA*RB point FFT of every column n2 < 243, four-step twiddle, W[n2*M1 + k1]
// ---------------------------------------------------------------------------------------
template <int RA, int RB>
__device__ inline void f2_pass1(const double *__restrict__ d, int N, int prefix, const float2 *__restrict__ tw1,
                                const float2 *__restrict__ T4, float2 *W, float2 *sm) {
    constexpr int M1 = RA * RB, P1 = M1 + 1;  // odd pitch: the column-strided stores are conflict free
    const int tid = threadIdx.x, nth = blockDim.x;
    const bool vec = (((uintptr_t)d >> 3) & 1u) == ((uint32_t)prefix & 1u);
    for (int c0 = 0; c0 < F2_M2; c0 += F2_TC) {
        const int nb = min(F2_TC, F2_M2 - c0);
        // stage 1: item (p < RB, column lc): radix RA over rows p + RB*t, Stockham twiddle W_M1^(p q)
        for (int item = tid; item < RB * F2_TC; item += nth) {
            const int lc = item & (F2_TC - 1), p = item / F2_TC;
            if (lc >= nb) continue;
            float2 a[RA];
#pragma unroll
            for (int t = 0; t < RA; t++) a[t] = f2_load_z(d, N, prefix, (p + RB * t) * F2_M2 + c0 + lc, vec);
            if (RA >= 4) {
                // outputs 1 and RA-1 feed the probed rows: sequential folds (see fold_step)
                float2 s1, sR;
                fold_pair<RA>(a, s1, sR);
                DftS<RA, 1, false>::run(a);
                a[1] = s1;
                a[RA - 1] = sR;
            } else {
                DftS<RA, 1, false>::run(a);
            }
            float2 *y = sm + lc * P1 + RA * p;
            y[0] = a[0];
#pragma unroll
            for (int q = 1; q < RA; q++) y[q] = cmul(a[q], __ldg(tw1 + p * q));
        }
        __syncthreads();
        // L2 prefetch of the next tile's samples (M1 rows x 512 B) while this tile finishes
        if (c0 + F2_TC < F2_M2) {
            for (int i = tid; i < M1 * 4; i += nth) {
                const int e = i >> 2, ln = i & 3;
                int ix = 2 * (e * F2_M2 + c0 + F2_TC) - prefix + 16 * ln;
                ix = min(max(ix, 0), N - 1);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(d + ix));
            }
        }
        // stage 2: item (q < RA, column lc), q fastest: radix RB over y[q + RA*t] -> k1 = q + RA*u
        for (int item = tid; item < RA * F2_TC; item += nth) {
            const int q = item % RA, lc = item / RA;
            if (lc >= nb) continue;
            float2 b[RB];
            const float2 *y = sm + lc * P1 + q;
#pragma unroll
            for (int t = 0; t < RB; t++) b[t] = y[RA * t];
            DftS<RB, 1, false>::run(b);
            const int base = (c0 + lc) * M1 + q;
#pragma unroll
            for (int u = 0; u < RB; u++) __stcg(W + base + RA * u, cmul(b[u], __ldg(T4 + base + RA * u)));
        }
        __syncthreads();
    }
}

// X[k] from Z[k] and Z[M-k]:  0.5 * ((Zk + conj Zm) - i w (Zk - conj Zm)),  w = exp(-2 pi i k / L)
__device__ __forceinline__ float2 f2_post(float2 Zk, float2 Zm, float2 w) {
    float2 sum = make_float2(Zk.x + Zm.x, Zk.y - Zm.y), dif = make_float2(Zk.x - Zm.x, Zk.y + Zm.y);
    float2 tt = cmul(w, dif);
    return make_float2(0.5f * (sum.x + tt.y), 0.5f * (sum.y - tt.x));
}
// Complex<f32>::norm() == hypotf; via f64 sqrt to stay correctly rounded (as fft_forward)
__device__ __forceinline__ uint32_t f2_key(float2 X) {
    double nr = sqrt((double)X.x * (double)X.x + (double)X.y * (double)X.y);
    return __float_as_uint((float)nr);
}

// ---------------------------------------------------------------------------------------
// pass 2: 243-point FFT of every row k1 + real-input post-process.
// Returns this thread's count of bins with X != 0 (fft.rs:249-252 stops at an exact zero).
// WRITE: also stores Xd / keys.
// ---------------------------------------------------------------------------------------
template <bool WRITE>
__device__ inline uint32_t f2_pass2(int M1, const float2 *__restrict__ tw2, const float2 *__restrict__ twL1,
                                    const float2 *__restrict__ twL2, const float2 *W, float2 *sm, float2 *Xd,
                                    uint32_t *keys) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int npairs = (M1 - 1) / 2;
    uint32_t nz = 0;
    for (int pr0 = 0; pr0 < npairs; pr0 += F2_PAIRS) {
        const int np = min(F2_PAIRS, npairs - pr0);
        // stage 1: item (p < 9, tile row lr): lr < 16 -> row pr0+1+lr, lr >= 16 -> its partner M1 - row
        for (int item = tid; item < 9 * 2 * F2_PAIRS; item += nth) {
            const int lr = item & (2 * F2_PAIRS - 1), p = item / (2 * F2_PAIRS), l = lr & (F2_PAIRS - 1);
            if (l >= np) continue;
            const int row = lr < F2_PAIRS ? pr0 + 1 + l : M1 - (pr0 + 1 + l);
            float2 a[27];
#pragma unroll
            for (int t = 0; t < 27; t++) a[t] = __ldcg(W + (p + 9 * t) * M1 + row);
            DftS<27, 1, false>::run(a);
            float2 *y = sm + lr * F2_M2 + 27 * p;
            y[0] = a[0];
#pragma unroll
            for (int q = 1; q < 27; q++) y[q] = cmul(a[q], __ldg(tw2 + p * q));
        }
        __syncthreads();
        // stage 2: item (pair pl, q < 27): rows k1 (element q) and M1-k1 (element 26-q) together
        for (int item = tid; item < F2_PAIRS * 27; item += nth) {
            const int q = item % 27, pl = item / 27;
            if (pl >= np) continue;
            float2 A[9], B[9];
            const float2 *ya = sm + pl * F2_M2 + q, *yb = sm + (pl + F2_PAIRS) * F2_M2 + (26 - q);
#pragma unroll
            for (int t = 0; t < 9; t++) {
                A[t] = ya[27 * t];
                B[t] = yb[27 * t];
            }
            DftS<9, 1, false>::run(A);
            DftS<9, 1, false>::run(B);
            const int k1 = pr0 + 1 + pl;
            const float2 w1 = __ldg(twL1 + k1);
#pragma unroll
            for (int u = 0; u < 9; u++) {
                const int k2 = q + 27 * u;
                const float2 w = cmul(w1, __ldg(twL2 + k2));            // exp(-2 pi i (k1 + M1 k2) / L)
                const float2 Xk = f2_post(A[u], B[8 - u], w);
                const float2 Xm = f2_post(B[8 - u], A[u], make_float2(-w.x, w.y));  // w_{M-k} = -conj(w_k)
                nz += (Xk.x != 0.f || Xk.y != 0.f) ? 1u : 0u;
                nz += (Xm.x != 0.f || Xm.y != 0.f) ? 1u : 0u;
                if (WRITE) {
                    const int i = k1 * F2_M2 + k2, im = (M1 - k1) * F2_M2 + (F2_M2 - 1 - k2);
                    Xd[i] = Xk;
                    keys[i] = f2_key(Xk);
                    Xd[im] = Xm;
                    keys[im] = f2_key(Xm);
                }
            }
        }
        __syncthreads();
    }
    // self-paired rows: k1 = 0 (partner (0, 243-k2)) and, for even M1, k1 = M1/2 (partner (M1/2, 242-k2))
    const int nself = (M1 % 2 == 0) ? 2 : 1;
    float2 *Zs = sm + 4 * F2_M2;
    for (int item = tid; item < 9 * nself; item += nth) {
        const int r = item / 9, p = item % 9, row = r ? M1 / 2 : 0;
        float2 a[27];
#pragma unroll
        for (int t = 0; t < 27; t++) a[t] = __ldcg(W + (p + 9 * t) * M1 + row);
        DftS<27, 1, false>::run(a);
        float2 *y = sm + r * F2_M2 + 27 * p;
        y[0] = a[0];
#pragma unroll
        for (int q = 1; q < 27; q++) y[q] = cmul(a[q], __ldg(tw2 + p * q));
    }
    __syncthreads();
    for (int item = tid; item < 27 * nself; item += nth) {
        const int r = item / 27, q = item % 27;
        float2 A[9];
#pragma unroll
        for (int t = 0; t < 9; t++) A[t] = sm[r * F2_M2 + q + 27 * t];
        DftS<9, 1, false>::run(A);
#pragma unroll
        for (int u = 0; u < 9; u++) Zs[r * F2_M2 + q + 27 * u] = A[u];
    }
    __syncthreads();
    for (int idx = tid; idx < F2_M2 * nself; idx += nth) {
        const int r = idx / F2_M2, k2 = idx - r * F2_M2, k1 = r ? M1 / 2 : 0;
        const int pk2 = r ? F2_M2 - 1 - k2 : (k2 ? F2_M2 - k2 : 0);
        const float2 Zk = Zs[r * F2_M2 + k2], Zm = Zs[r * F2_M2 + pk2];
        const float2 w = cmul(__ldg(twL1 + k1), __ldg(twL2 + k2));
        const float2 X = f2_post(Zk, Zm, w);
        nz += (X.x != 0.f || X.y != 0.f) ? 1u : 0u;
        if (WRITE) {
            Xd[k1 * F2_M2 + k2] = X;
            keys[k1 * F2_M2 + k2] = f2_key(X);
        }
        if (idx == 0) {
            const float2 XM = make_float2(Zk.x - Zk.y, 0.f);  // Nyquist bin X[M] = Re Z0 - Im Z0
            nz += XM.x != 0.f ? 1u : 0u;
            if (WRITE) {
                Xd[M1 * F2_M2] = XM;
                keys[M1 * F2_M2] = f2_key(XM);
            }
        }
    }
    __syncthreads();
    return nz;
}

// ---------------------------------------------------------------------------------------
// PROBE: the same transform restricted to the 2*RB rows k1 = 1 + RA*u and their partners
// M1 - k1 = (RA-1) + RA*(RB-1-u) (u < RB), i.e. 2*RB*243 of the M bins, bit-identical to what the
// full transform computes for them.  Used to prove "at least c nonzero bins" (Auto pruning) at
// roughly a third of the full cost: pass-1 stage 2 runs for q in {1, RA-1} only, pass 2 for 2*RB
// rows only, and the intermediate is 1/8 of W:  Wp[n2 * 2RB + lr],  lr < RB <-> k1 = 1 + RA*lr,
// lr >= RB <-> k1 = M1 - (1 + RA*(lr - RB)).
// ---------------------------------------------------------------------------------------
template <int RA, int RB>
__device__ inline void f2_probe_pass1(const double *__restrict__ d, int N, int prefix, const float2 *__restrict__ tw1,
                                      const float2 *__restrict__ T4, float2 *Wp, float2 *sm) {
    static_assert(RA >= 4, "probe needs two distinct output families");
    static_assert(F2_M2 * (2 * RB + 1) <= F2_SMEM_F2, "all columns' stage-1 outputs must fit the tile buffer");
    constexpr int M1 = RA * RB, P1 = 2 * RB + 1;  // only y[q + RA*t] for the two q is kept: [column][2][RB]
    const int tid = threadIdx.x, nth = blockDim.x;
    const bool vec = (((uintptr_t)d >> 3) & 1u) == ((uint32_t)prefix & 1u);
    // stage 1 for ALL columns (no tile barriers: every thread streams through its items, so the
    // sample loads of one item overlap the butterflies of the previous one across warps)
    for (int item = tid; item < RB * F2_M2; item += nth) {
        const int p = item / F2_M2, c = item - p * F2_M2;
        float2 a[RA];
#pragma unroll
        for (int t = 0; t < RA; t++) a[t] = f2_load_z(d, N, prefix, (p + RB * t) * F2_M2 + c, vec);
        float2 s1, sR;
        fold_pair<RA>(a, s1, sR);  // only outputs 1 and RA-1 are needed
        float2 *y = sm + c * P1;
        y[p] = cmul(s1, __ldg(tw1 + p));
        y[RB + p] = cmul(sR, __ldg(tw1 + p * (RA - 1)));
    }
    __syncthreads();
    for (int item = tid; item < 2 * F2_M2; item += nth) {
        const int fam = item & 1, c = item >> 1;  // fam 0: q = 1, fam 1: q = RA-1
        float2 b[RB];
        const float2 *y = sm + c * P1 + fam * RB;
#pragma unroll
        for (int t = 0; t < RB; t++) b[t] = y[t];
        DftS<RB, 1, false>::run(b);
        const int q = fam ? RA - 1 : 1;
#pragma unroll
        for (int u = 0; u < RB; u++) {
            const int k1 = q + RA * u, lr = fam ? RB + (RB - 1 - u) : u;
            __stcg(Wp + c * (2 * RB) + lr, cmul(b[u], __ldg(T4 + c * M1 + k1)));
        }
    }
    __syncthreads();
}

// count of nonzero bins among the probed rows (this thread's share)
__device__ inline uint32_t f2_probe_pass2(int M1, int RA, const float2 *__restrict__ tw2, const float2 *__restrict__ twL1,
                                          const float2 *__restrict__ twL2, const float2 *Wp, float2 *sm) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int RB = M1 / RA, rows = 2 * RB;  // <= 36 rows: one tile (36 * 243 float2 fits the buffer)
    uint32_t nz = 0;
    // stage 1: item (p < 9, compact row lr): lr < RB -> k1 = 1 + RA*lr, lr >= RB -> its partner
    for (int item = tid; item < 9 * rows; item += nth) {
        const int p = item / rows, lr = item - p * rows;
        float2 a[27];
#pragma unroll
        for (int t = 0; t < 27; t++) a[t] = __ldcg(Wp + (p + 9 * t) * rows + lr);
        DftS<27, 1, false>::run(a);
        float2 *y = sm + lr * F2_M2 + 27 * p;
        y[0] = a[0];
#pragma unroll
        for (int q = 1; q < 27; q++) y[q] = cmul(a[q], __ldg(tw2 + p * q));
    }
    __syncthreads();
    // stage 2: item (pair pl < RB, q < 27): rows k1 (element q) and M1-k1 (element 26-q) together
    for (int item = tid; item < RB * 27; item += nth) {
        const int q = item % 27, pl = item / 27;
        float2 A[9], B[9];
        const float2 *ya = sm + pl * F2_M2 + q, *yb = sm + (pl + RB) * F2_M2 + (26 - q);
#pragma unroll
        for (int t = 0; t < 9; t++) {
            A[t] = ya[27 * t];
            B[t] = yb[27 * t];
        }
        DftS<9, 1, false>::run(A);
        DftS<9, 1, false>::run(B);
        const int k1 = 1 + RA * pl;
        const float2 w1 = __ldg(twL1 + k1);
#pragma unroll
        for (int u = 0; u < 9; u++) {
            const int k2 = q + 27 * u;
            const float2 w = cmul(w1, __ldg(twL2 + k2));
            const float2 Xk = f2_post(A[u], B[8 - u], w);
            const float2 Xm = f2_post(B[8 - u], A[u], make_float2(-w.x, w.y));
            nz += (Xk.x != 0.f || Xk.y != 0.f) ? 1u : 0u;
            nz += (Xm.x != 0.f || Xm.y != 0.f) ? 1u : 0u;
        }
    }
    __syncthreads();
    return nz;
}

// ---------------------------------------------------------------------------------------
// INVERSE of the first `c` list entries (the refinement loop's per-iteration transform):
// Epi(j, value) is called once for every time index j < L with the unnormalised real output.
// Mirror image of the forward engine with conjugated roots:
//   pass 1'  rows k1:    sparse list -> zero-filled smem tile (scatter coefficients of
//            fft_prepare_entries) -> registers (radix 27) -> smem -> registers (radix 9) -> W[k1][n2]
//   pass 2'  columns n2: W * conj(T4T) -> registers (radix RA) -> smem -> registers (radix RB)
//            -> epilogue on samples 2n, 2n+1 (n = n1*243 + n2), 32 consecutive columns per warp.
// Written for any block size (item loops); tile sizes assume <= 512 threads hold one pass-1' item each.
// ---------------------------------------------------------------------------------------
constexpr int F2I_TR = 56;   // rows per pass-1' tile: 9 * 56 = 504 radix-27 items
constexpr int F2I_TC = 28;   // columns per pass-2' tile: 18 * 28 = 504 radix-16 items
constexpr int F2I_SMEM_F2 = F2I_TR * F2_M2;  // 13,608 float2 = 108,864 B  (>= 28 * 289)

template <int RA, int RB, class Epi>
__device__ inline void f2_inv_pass2(const float2 *__restrict__ tw1, const float2 *__restrict__ T4T, const float2 *W,
                                    float2 *sm, Epi epi) {
    constexpr int M1 = RA * RB, P1 = M1 + 1;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int c0 = 0; c0 < F2_M2; c0 += F2I_TC) {
        const int nb = min(F2I_TC, F2_M2 - c0);
        // stage 1: item (p < RB, column lc), lc fastest: radix RA over rows k1 = p + RB*t
        for (int item = tid; item < RB * nb; item += nth) {
            const int p = item / nb, lc = item - p * nb;
            float2 a[RA];
#pragma unroll
            for (int t = 0; t < RA; t++) {
                const int i = (p + RB * t) * F2_M2 + c0 + lc;
                a[t] = cmulc(__ldcg(W + i), __ldg(T4T + i));  // four-step twiddle exp(+2 pi i k1 n2 / M)
            }
            DftS<RA, 1, true>::run(a);
            float2 *y = sm + lc * P1 + RA * p;
            y[0] = a[0];
#pragma unroll
            for (int q = 1; q < RA; q++) y[q] = cmulc(a[q], __ldg(tw1 + p * q));
        }
        __syncthreads();
        // stage 2: item (q < RA, column lc), lc fastest: radix RB over y[q + RA*t] -> n1 = q + RA*u
        for (int item = tid; item < RA * nb; item += nth) {
            const int q = item / nb, lc = item - q * nb;
            float2 b[RB];
            const float2 *y = sm + lc * P1 + q;
#pragma unroll
            for (int t = 0; t < RB; t++) b[t] = y[RA * t];
            DftS<RB, 1, true>::run(b);
#pragma unroll
            for (int u = 0; u < RB; u++) {
                const uint32_t n = (uint32_t)((q + RA * u) * F2_M2 + c0 + lc);
                epi(2 * n, b[u].x);
                epi(2 * n + 1, b[u].y);
            }
        }
        __syncthreads();
    }
}

template <class Epi>
__device__ inline void f2_inverse(const FftGeom &g, FftWs ws, uint32_t c, float2 *sm, Epi epi) {
    const int M1 = (int)g.M1;
    const int tid = threadIdx.x, nth = blockDim.x;
    const float2 *__restrict__ tw2 = g.tw2;
    // ---- pass 1': row tiles built from the sparse list
    for (int r0 = 0; r0 < M1; r0 += F2I_TR) {
        const int nb = min(F2I_TR, M1 - r0);
        for (int i = tid; i < nb * F2_M2; i += nth) sm[i] = make_float2(0.f, 0.f);
        __syncthreads();
        for (uint32_t r = tid; r < c; r += nth) {
            const uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;  // overwritten by a later aliased entry (fft.rs:411-420)
            const uint32_t l = ws.locD[r];
            if (l != 0xFFFFFFFFu) {
                const uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < (uint32_t)nb) sm[f * F2_M2 + (l & 0xFFFFu)] = ws.cD[r];
            }
        }
        __syncthreads();
        for (uint32_t r = tid; r < c; r += nth) {
            const uint32_t ov = ws.ovr[r];
            if (ov > r && ov < c) continue;
            const uint32_t l = ws.locM[r];
            if (l != 0xFFFFFFFFu) {
                const uint32_t f = (l >> 16) - (uint32_t)r0;
                if (f < (uint32_t)nb) {
                    float2 *q = &sm[f * F2_M2 + (l & 0xFFFFu)];
                    *q = cadd(*q, ws.cM[r]);
                }
            }
        }
        __syncthreads();
        // stage 1 (in place: every item is read into registers before any is written back)
        for (int base = 0; base < 9 * nb; base += nth) {
            const int item = base + tid;
            const bool on = item < 9 * nb;
            const int p = on ? item / nb : 0, lr = on ? item - p * nb : 0;
            float2 a[27];
            if (on) {
#pragma unroll
                for (int t = 0; t < 27; t++) a[t] = sm[lr * F2_M2 + p + 9 * t];
                DftS<27, 1, true>::run(a);
            }
            __syncthreads();
            if (on) {
                float2 *y = sm + lr * F2_M2 + 27 * p;
                y[0] = a[0];
#pragma unroll
                for (int q = 1; q < 27; q++) y[q] = cmulc(a[q], __ldg(tw2 + p * q));
            }
            __syncthreads();
        }
        // stage 2: item (row lr, q < 27), q fastest: radix 9 over y[q + 27 t] -> n2 = q + 27 u
        for (int item = tid; item < nb * 27; item += nth) {
            const int lr = item / 27, q = item - lr * 27;
            float2 b[9];
#pragma unroll
            for (int t = 0; t < 9; t++) b[t] = sm[lr * F2_M2 + q + 27 * t];
            DftS<9, 1, true>::run(b);
            float2 *o = ws.W + (size_t)(r0 + lr) * F2_M2 + q;
#pragma unroll
            for (int u = 0; u < 9; u++) __stcg(o + 27 * u, b[u]);
        }
        __syncthreads();
    }
    // ---- pass 2': column tiles + epilogue
    switch (M1) {
        case 288: f2_inv_pass2<16, 18>(g.tw1, g.T4T, ws.W, sm, epi); break;
        case 144: f2_inv_pass2<16, 9>(g.tw1, g.T4T, ws.W, sm, epi); break;
        case 72: f2_inv_pass2<8, 9>(g.tw1, g.T4T, ws.W, sm, epi); break;
        case 36: f2_inv_pass2<4, 9>(g.tw1, g.T4T, ws.W, sm, epi); break;
        default: f2_inv_pass2<2, 9>(g.tw1, g.T4T, ws.W, sm, epi); break;
    }
}

__device__ inline bool f2_supported(const FftGeom &g) {
    return g.real && g.M2 == (uint32_t)F2_M2 && g.T4 != nullptr &&
           (g.M1 == 288u || g.M1 == 144u || g.M1 == 72u || g.M1 == 36u || g.M1 == 18u);
}

__device__ inline void f2_forward_pass1(const double *__restrict__ d, int N, int prefix, const FftGeom &g, float2 *W,
                                        float2 *sm) {
    switch (g.M1) {
        case 288: f2_pass1<16, 18>(d, N, prefix, g.tw1, g.T4, W, sm); break;
        case 144: f2_pass1<16, 9>(d, N, prefix, g.tw1, g.T4, W, sm); break;
        case 72: f2_pass1<8, 9>(d, N, prefix, g.tw1, g.T4, W, sm); break;
        case 36: f2_pass1<4, 9>(d, N, prefix, g.tw1, g.T4, W, sm); break;
        default: f2_pass1<2, 9>(d, N, prefix, g.tw1, g.T4, W, sm); break;
    }
}

// probe (see above): this thread's count of nonzero bins among 2*RB*243 probed ones; 0 when the
// geometry has no probe (M1 = 18)
__device__ inline uint32_t f2_probe(const double *__restrict__ d, int N, int prefix, const FftGeom &g, float2 *W, float2 *sm) {
    int RA;
    switch (g.M1) {
        case 288: f2_probe_pass1<16, 18>(d, N, prefix, g.tw1, g.T4, W, sm); RA = 16; break;
        case 144: f2_probe_pass1<16, 9>(d, N, prefix, g.tw1, g.T4, W, sm); RA = 16; break;
        case 72: f2_probe_pass1<8, 9>(d, N, prefix, g.tw1, g.T4, W, sm); RA = 8; break;
        case 36: f2_probe_pass1<4, 9>(d, N, prefix, g.tw1, g.T4, W, sm); RA = 4; break;
        default: return 0u;
    }
    return f2_probe_pass2((int)g.M1, RA, g.tw2, g.twL1, g.twL2, W, sm);
}

// ---------------------------------------------------------------------------------------
// PROBE from a fold: k_front (front.cuh) accumulates the two folds (A, B) of every slot
// m = p*243 + c < RB*243 as one float4  fold[m] = (A.x, A.y, B.x, B.y) in shared memory and runs
// the rest of the probe at the end of the frame.  What remains is fold_out,
// the Stockham twiddle, pass-1 stage 2 for the two families and pass 2 for the 2*RB probed rows.
// ---------------------------------------------------------------------------------------
template <int RA, int RB>
__device__ inline void f2_fold_stage2(const float4 *__restrict__ fold, const float2 *__restrict__ tw1,
                                      const float2 *__restrict__ T4, float2 *Wp) {
    constexpr int M1 = RA * RB;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int item = tid; item < 2 * F2_M2; item += nth) {
        const int fam = item & 1, c = item >> 1;  // fam 0: q = 1, fam 1: q = RA-1
        const int q = fam ? RA - 1 : 1;
        float2 b[RB];
#pragma unroll
        for (int t = 0; t < RB; t++) {
            float2 s1, sR;
            fold_out(fold[t * F2_M2 + c], s1, sR);  // shared memory (k_front) or global
            b[t] = cmul(fam ? sR : s1, __ldg(tw1 + t * q));
        }
        DftS<RB, 1, false>::run(b);
#pragma unroll
        for (int u = 0; u < RB; u++) {
            const int k1 = q + RA * u, lr = fam ? RB + (RB - 1 - u) : u;
            __stcg(Wp + c * (2 * RB) + lr, cmul(b[u], __ldg(T4 + c * M1 + k1)));
        }
    }
    __syncthreads();
}
__device__ inline uint32_t f2_probe_from_fold(const float4 *__restrict__ fold, const FftGeom &g, float2 *W, float2 *sm) {
    int RA;
    switch (g.M1) {
        case 288: f2_fold_stage2<16, 18>(fold, g.tw1, g.T4, W); RA = 16; break;
        case 144: f2_fold_stage2<16, 9>(fold, g.tw1, g.T4, W); RA = 16; break;
        case 72: f2_fold_stage2<8, 9>(fold, g.tw1, g.T4, W); RA = 8; break;
        case 36: f2_fold_stage2<4, 9>(fold, g.tw1, g.T4, W); RA = 4; break;
        default: return 0u;
    }
    return f2_probe_pass2((int)g.M1, RA, g.tw2, g.twL1, g.twL2, W, sm);
}

// ---------------------------------------------------------------------------------------
// SMALL PROBE from a fold (k_probe): the pruning rule only needs "at least c nonzero bins" with
// c <= max_freq (fft.rs:249-252), so `np` row pairs (k1 = 1 + RA*pl and its partner M1 - k1, pl < np:
// 2 * np * 243 bins) are enough whenever they are all nonzero -- the usual case.  Same expressions as
// f2_fold_stage2 / f2_probe_pass2 on those rows (bit-identical bins); everything after the fold stays in
// shared memory (smW, smY: 2 * np * 243 float2 each).  Returns this thread's count of nonzero bins.
// ---------------------------------------------------------------------------------------
constexpr int F2_PROBE_NP = 4;  // pairs at most: 8 rows
template <int RA, int RB>
__device__ inline uint32_t f2_probe_small_t(const float4 *__restrict__ fold, const FftGeom &g, int np, float2 *smW,
                                            float2 *smY) {
    constexpr int M1 = RA * RB;
    const int tid = threadIdx.x, nth = blockDim.x, rows = 2 * np;
    // pass-1 stage 2 of the two families; only the rows of the first np pairs are kept
    for (int item = tid; item < 2 * F2_M2; item += nth) {
        const int fam = item & 1, c = item >> 1;  // fam 0: q = 1, fam 1: q = RA-1
        const int q = fam ? RA - 1 : 1;
        float2 b[RB];
#pragma unroll
        for (int t = 0; t < RB; t++) {
            float2 s1, sR;
            fold_out(__ldcg(fold + t * F2_M2 + c), s1, sR);
            b[t] = cmul(fam ? sR : s1, __ldg(g.tw1 + t * q));
        }
        DftS<RB, 1, false>::run(b);
#pragma unroll
        for (int u = 0; u < RB; u++) {
            const int pl = fam ? RB - 1 - u : u;  // the pair this row belongs to
            if (pl < np) {
                const int k1 = q + RA * u, lr = fam ? np + pl : pl;
                smW[c * rows + lr] = cmul(b[u], __ldg(g.T4 + c * M1 + k1));
            }
        }
    }
    __syncthreads();
    // pass 2, stage 1: item (p < 9, compact row lr)
    for (int item = tid; item < 9 * rows; item += nth) {
        const int p = item / rows, lr = item - p * rows;
        float2 a[27];
#pragma unroll
        for (int t = 0; t < 27; t++) a[t] = smW[(p + 9 * t) * rows + lr];
        DftS<27, 1, false>::run(a);
        float2 *y = smY + lr * F2_M2 + 27 * p;
        y[0] = a[0];
#pragma unroll
        for (int q = 1; q < 27; q++) y[q] = cmul(a[q], __ldg(g.tw2 + p * q));
    }
    __syncthreads();
    // pass 2, stage 2: item (pair pl < np, q < 27): rows k1 (element q) and M1-k1 (element 26-q) together
    uint32_t nz = 0;
    for (int item = tid; item < np * 27; item += nth) {
        const int q = item % 27, pl = item / 27;
        float2 A[9], B[9];
        const float2 *ya = smY + pl * F2_M2 + q, *yb = smY + (pl + np) * F2_M2 + (26 - q);
#pragma unroll
        for (int t = 0; t < 9; t++) {
            A[t] = ya[27 * t];
            B[t] = yb[27 * t];
        }
        DftS<9, 1, false>::run(A);
        DftS<9, 1, false>::run(B);
        const int k1 = 1 + RA * pl;
        const float2 w1 = __ldg(g.twL1 + k1);
#pragma unroll
        for (int u = 0; u < 9; u++) {
            const int k2 = q + 27 * u;
            const float2 w = cmul(w1, __ldg(g.twL2 + k2));
            const float2 Xk = f2_post(A[u], B[8 - u], w);
            const float2 Xm = f2_post(B[8 - u], A[u], make_float2(-w.x, w.y));
            nz += (Xk.x != 0.f || Xk.y != 0.f) ? 1u : 0u;
            nz += (Xm.x != 0.f || Xm.y != 0.f) ? 1u : 0u;
        }
    }
    __syncthreads();
    return nz;
}
__device__ inline uint32_t f2_probe_small(const float4 *__restrict__ fold, const FftGeom &g, int np, float2 *smW, float2 *smY) {
    switch (g.M1) {
        case 288: return f2_probe_small_t<16, 18>(fold, g, np, smW, smY);
        case 144: return f2_probe_small_t<16, 9>(fold, g, np, smW, smY);
        case 72: return f2_probe_small_t<8, 9>(fold, g, np, smW, smY);
        case 36: return f2_probe_small_t<4, 9>(fold, g, np, smW, smY);
        default: return 0u;
    }
}

}  // namespace atsc
