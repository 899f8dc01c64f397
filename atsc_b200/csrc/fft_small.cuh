// fft_small.cuh -- FFT compressor for short frames (transform length L <= 1152: every frame of
// <= 1024 samples, i.e. the tail frames of a series).  Same routine as fft_frame (kernels.cu):
//   FFT::compress_bounded  fft.rs:288-362     FFT::fft_trim  fft.rs:231-257
//   get_mirrored_freqs     fft.rs:401-422     FFT::round     fft.rs:208-218
// but sized for the job: one 128-thread CTA per frame, everything in shared memory, and the
// transforms evaluated directly -- the forward one as a dense DFT (L * L/2 terms), the inverse of
// every refinement iteration as a SPARSE sum over the <= 100 kept bins, which for these lengths is
// cheaper than any FFT and needs no block-wide staging.  The big-frame engine (fft.cuh / fft2.cuh)
// spends ~100 us of barrier latency on such a frame; this path a few tens, eight frames per SM.
#pragma once
#include "fft.cuh"

namespace atsc {

constexpr int FS_THREADS = 128;
constexpr int FS_LMAX = 1152;               // next_size(1024)
constexpr int FS_BMAX = FS_LMAX / 2 + 1;    // half-spectrum bins
constexpr int FS_NJ = FS_LMAX / FS_THREADS;                       // outputs per thread (9)
constexpr int FS_KMAX = 128;                // >= max_freq schedule maximum for N <= 1024 (10 + 17*5 + 5*1 = 100)

// Sums run in f64 (the f32 operands are exact in f64, so every bin / sample is the correctly rounded
// value of the exact DFT of the f32 inputs: inside the f32 noise band of any f32 FFT, the
// reference's included, even for DC-dominated frames where a naive f32 sum would not be).
// Shared-memory arrays are carved for the longest transform of the launch (lmax), so a wave of
// 512-sample tails runs seven CTAs per SM instead of three.
struct FsSmem {
    double2 *w;              // [lmax]  exp(-2 pi i j / L)
    double2 *Y;              // [lmax]  level-1 result of the forward transform
    double *x;               // [lmax]  padded samples rounded to f32 (fft.rs:221-228), widened
    unsigned long long *S;   // [pow2 >= lmax/2 + 1] composite sort keys (|z| bits << 20 | inverted bin)
    float2 *X;               // [lmax/2 + 2] half spectrum
    FftEntry *e;             // [FS_KMAX] kept entries, descending |z|
    double *red;             // [40]
};
__host__ __device__ inline uint32_t fs_sort_len(uint32_t lmax) {
    uint32_t P = 2;
    while (P < lmax / 2 + 1) P <<= 1;
    return P;
}
__host__ __device__ inline size_t fs_smem_bytes(uint32_t lmax) {
    return (size_t)lmax * (16 + 16 + 8) + (size_t)fs_sort_len(lmax) * 8 + ((size_t)lmax / 2 + 2) * 8 +
           (size_t)FS_KMAX * sizeof(FftEntry) + 40 * 8 + 64;
}
__device__ inline FsSmem fs_carve(unsigned char *base, uint32_t lmax) {
    FsSmem m;
    m.w = reinterpret_cast<double2 *>(base);
    m.Y = m.w + lmax;
    m.x = reinterpret_cast<double *>(m.Y + lmax);
    m.S = reinterpret_cast<unsigned long long *>(m.x + lmax);
    m.X = reinterpret_cast<float2 *>(m.S + fs_sort_len(lmax));
    m.red = reinterpret_cast<double *>(m.X + lmax / 2 + 2);
    m.e = reinterpret_cast<FftEntry *>(m.red + 40);
    return m;
}

__device__ inline double fs_block_sum(double v, double *scratch) { return block_sum(v, scratch); }

// all threads of the (128-thread) CTA call; result fields of fw written by thread 0
__device__ inline void fft_small_frame(const double *__restrict__ d, FrameWork *fw, const FftGeom *__restrict__ geoms,
                                       FftEntry *list, double max_err, const FsSmem *sm) {
    const uint32_t N = fw->len, T = blockDim.x, t = threadIdx.x;
    const bool bounded = fw->bounded != 0;
    uint32_t *sh = (uint32_t *)sm->red;
    if (fw->f32_const) {
        // fft.rs:289-292 "Same max and min": no frequencies, error None -> 0.0
        if (t == 0) {
            fw->fft_count = 0;
            fw->fft_err = 0.0;
            fw->fft_size = fft_payload_size(0, 0);
            fw->fft_iters = 0;
            fw->fft_tie = 0;
            fw->fft_valid = 1;
        }
        return;
    }
    const float vminf = (float)fw->vmin, vmaxf = (float)fw->vmax;
    const uint32_t mf = (3 >= N / 100) ? 3 : N / 100;
    const uint32_t hstep = max(mf / 2, 1u), tstep = max(mf / 10, 1u);
    const uint32_t kmax = bounded ? mf + 17 * hstep + 5 * tstep : mf;
    const int gi = fw->geom;
    const uint32_t L = gi >= 0 ? geoms[gi].L : N;
    const uint32_t Bn = L / 2 + 1;
    const uint32_t prefix = (gi >= 0 && bounded && N >= 128) ? (L - N) / 2 : 0u;
    __syncthreads();
    // ---- stage the padded frame (gibbs sizing, fft.rs:184-204) and the roots of unity
    for (uint32_t j = t; j < L; j += T) {
        uint32_t ix = j < prefix ? 0u : j - prefix;
        if (ix >= N) ix = N - 1;
        sm->x[j] = (double)(float)d[ix];
    }
    // roots of unity in f64 (sincospi of an exact rational argument): the recurrences below multiply
    // them hundreds of times, which f32-rounded table values would not survive
    for (uint32_t j = t; j < L; j += T) {
        double sn, cs;
        sincospi(2.0 * (double)j / (double)L, &sn, &cs);
        sm->w[j] = make_double2(cs, -sn);
    }
    __syncthreads();
    // ---- forward DFT, bins 0 .. L/2, and their sort keys (Complex<f32>::norm() == hypotf)
    // Forward transform in two levels, L = L1 * L2 (L1 = largest divisor <= sqrt(L); 1 for a prime L,
    // which only an unpadded frame of < 128 samples can have): L1-point DFTs over n1, the four-step
    // twiddle, then L2-point DFTs over n2 -- L * (L1 + L2/2) terms instead of L * L/2, each sum
    // carried in f64 with exact roots.
    uint32_t L1 = 1;
    for (uint32_t q = 2; q * q <= L; q++)
        if (L % q == 0) L1 = q;
    const uint32_t L2 = L / L1;
    uint32_t nzl = 0;
    // level 1: Y[k1][n2] = W_L^(k1 n2) * sum_n1 x[n1 L2 + n2] W_L1^(n1 k1)
    for (uint32_t o = t; o < L; o += T) {
        const uint32_t k1 = o / L2, n2 = o - k1 * L2;
        const uint32_t stepi = (k1 * L2) % L;  // W_L1^(k1) = w[k1 L2]
        double ar = 0.0, ai = 0.0;
        uint32_t idx = 0;
        for (uint32_t n1 = 0; n1 < L1; n1++) {
            const double2 w = sm->w[idx];
            const double xv = sm->x[n1 * L2 + n2];
            ar = fma(xv, w.x, ar);
            ai = fma(xv, w.y, ai);
            idx += stepi;
            if (idx >= L) idx -= L;
        }
        const double2 tw = sm->w[k1 * n2];
        sm->Y[o] = make_double2(fma(ar, tw.x, -ai * tw.y), fma(ar, tw.y, ai * tw.x));
    }
    __syncthreads();
    // level 2: X[k1 + L1 k2] = sum_n2 Y[k1][n2] W_L2^(n2 k2), bins k <= L/2 only
    for (uint32_t o = t; o < L; o += T) {
        const uint32_t k2 = o / L1, k1 = o - k2 * L1, k = k1 + L1 * k2;  // o == k: bins in natural order
        if (k >= Bn) continue;
        const uint32_t stepi = (k2 * L1) % L;  // W_L2^(k2) = w[k2 L1]
        const double2 *Yr = sm->Y + k1 * L2;
        double ar = 0.0, ai = 0.0;
        uint32_t idx = 0;
        for (uint32_t n2 = 0; n2 < L2; n2++) {
            const double2 w = sm->w[idx], y = Yr[n2];
            ar = fma(y.x, w.x, fma(-y.y, w.y, ar));
            ai = fma(y.x, w.y, fma(y.y, w.x, ai));
            idx += stepi;
            if (idx >= L) idx -= L;
        }
        float2 X = make_float2((float)ar, (float)ai);
        if (k == 0 || 2 * k == L) X.y = 0.f;  // purely real bins of a real signal
        sm->X[k] = X;
    }
    __syncthreads();
    uint32_t P = 2;
    while (P < Bn) P <<= 1;
    for (uint32_t k = t; k < P; k += T) {
        unsigned long long comp = 0ull;
        if (k < Bn) {
            const float2 X = sm->X[k];
            const double nr = sqrt((double)X.x * (double)X.x + (double)X.y * (double)X.y);
            const uint32_t key = __float_as_uint((float)nr);  // Complex<f32>::norm() == hypotf, correctly rounded
            if (key != 0u) {
                comp = fft_composite(key, k);
                nzl++;
            }
        }
        sm->S[k] = comp;
    }
    const uint32_t nz = block_sum_u32(nzl, sh);
    // a later candidate can only lose to FFT on size; FFT wins ties (frame/mod.rs:77,104,141)
    uint32_t bound = 0xFFFFFFFFu;
    if (bounded && fw->comp == C_AUTO && fw->forced == 0xFF) {
        if (fw->poly_valid == 1 && fw->poly_err <= max_err) bound = min(bound, fw->poly_size);
        if (fw->need_rle) bound = min(bound, fw->rle_valid == 1 ? fw->rle_size : rle_upper_bound(fw));  // RLE's error is 0: it always passes
    }
    const uint32_t kcap = min(min(kmax, fw->fft_list_cap), (uint32_t)FS_KMAX);
    if (bound != 0xFFFFFFFFu) {
        const uint32_t c1 = min(min(mf, nz), kcap);
        if (fft_payload_size(c1, min(c1, 251u)) > bound) {
            if (t == 0) {
                fw->fft_count = c1;
                fw->fft_err = max_err + 1.0;
                fw->fft_size = 0;
                fw->fft_iters = 1;
                fw->fft_tie = 0;
                fw->fft_valid = 2;
            }
            return;
        }
    }
    // ---- order every nonzero bin: |z| descending, then lower bin (fft.rs:231-257; the pop order
    // among equal norms is unspecified in the reference -> near-tie flag at such a cut)
    for (uint32_t k2 = 2; k2 <= P; k2 <<= 1) {
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = t; i < P; i += T) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sm->S[i], b = sm->S[ixj];
                    const bool desc = (i & k2) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        sm->S[i] = b;
                        sm->S[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    const uint32_t K = min(kcap, nz);
    for (uint32_t r = t; r < K; r += T) {
        const uint32_t bin = 0xFFFFFu - (uint32_t)(sm->S[r] & 0xFFFFFull);
        FftEntry e;
        e.bin = bin;
        e.re = sm->X[bin].x;
        e.im = sm->X[bin].y;
        sm->e[r] = e;
        list[r] = e;
    }
    __syncthreads();
    auto tie_at = [&](uint32_t c) { return c > 0 && c < nz && (sm->S[c - 1] >> 20) == (sm->S[c] >> 20); };
    uint32_t cutmask = 0;
    {
        uint32_t jump = 0;
        for (int it = 1; it <= FFT_SCHED; it++) {
            const uint32_t c = min(mf + jump, K);
            if (tie_at(c)) cutmask |= 1u << (it - 1);
            jump += it <= 17 ? hstep : tstep;
            if (!bounded || mf + jump > K) break;
        }
    }
    auto nsmall = [&](uint32_t c) -> uint32_t {
        uint32_t loc = 0;
        for (uint32_t r = t; r < c; r += T) loc += sm->e[r].bin < 251u;
        return block_sum_u32(loc, sh);
    };
    if (!bounded) {
        // FFT::compress (fft.rs:366-388): max(3, n/100) frequencies, no refinement
        const uint32_t ns = nsmall(K);
        if (t == 0) {
            fw->fft_count = K;
            fw->fft_err = 0.0;
            fw->fft_size = fft_payload_size(K, ns);
            fw->fft_iters = 0;
            fw->fft_tie = (cutmask & 1u) ? TIE_FFT_TOPK : 0;
            fw->fft_valid = 1;
        }
        return;
    }
    // ---- refinement loop (fft.rs:334-353)
    const float Lf = (float)L;
    const uint32_t magic = (uint32_t)((0x100000000ull + L - 1) / L);  // (a / L) for a < 2^20 * ... : a = bin * j < 2^21
    // Sparse inverse: thread t owns the FS_NJ consecutive outputs j0 .. j0+FS_NJ-1; per kept bin one
    // table lookup gives the phase at j0 and a (broadcast) one the per-sample step.
    const uint32_t per = (L + T - 1) / T, j0 = t * per;
    auto evaluate = [&](uint32_t c) -> double {
        double v[FS_NJ];
#pragma unroll
        for (int s2 = 0; s2 < FS_NJ; s2++) v[s2] = 0.0;
        for (uint32_t r = 0; r < c; r++) {
            const FftEntry e = sm->e[r];
            const double re = (double)e.re, im = (double)e.im;
            if (e.bin == 0) {
#pragma unroll
                for (int s2 = 0; s2 < FS_NJ; s2++) v[s2] += re;
            } else if (2 * e.bin == L) {
#pragma unroll
                for (int s2 = 0; s2 < FS_NJ; s2++) v[s2] += ((j0 + s2) & 1u) ? -re : re;
            } else {
                const uint32_t a = e.bin * (j0 < L ? j0 : 0u);
                const double2 ws = sm->w[a - __umulhi(a, magic) * L], wb = sm->w[e.bin];
                // exp(+i phi) = conj of the table entries
                double pr = ws.x, pi = -ws.y;
                const double br = wb.x, bi = -wb.y;
#pragma unroll
                for (int s2 = 0; s2 < FS_NJ; s2++) {
                    v[s2] += 2.0 * fma(re, pr, -im * pi);
                    const double nr = fma(pr, br, -pi * bi);
                    pi = fma(pr, bi, pi * br);
                    pr = nr;
                }
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < FS_NJ; s2++) {
            const uint32_t j = j0 + s2;
            if (s2 < (int)per && j < L) {
                uint32_t ix = j < prefix ? 0u : j - prefix;  // gibbs padding replicates the edge samples
                if (ix >= N) ix = N - 1;
                const double out = fft_round_fast(__fdiv_rn((float)v[s2], Lf), vminf, vmaxf);
                acc += mape_term(out, d[ix]);
            }
        }
        const double s = block_sum(acc, sm->red);
        return __ddiv_rn(s, (double)L);
    };
    const int E = rust_as_i32(max_err * 1000.0);
    double cur = max_err + 1.0, prev_err = 0.0;
    uint32_t jump = 0, it = 0, c = 0, prev_c = 0xFFFFFFFFu;
    bool pruned = false, tie = false, tie_topk = false;
    while (E < rust_as_i32(cur * 1000.0)) {
        it++;
        c = min(mf + jump, K);
        if (bound != 0xFFFFFFFFu) {
            const uint32_t ns = nsmall(c);
            if (fft_payload_size(c, ns) > bound) {
                pruned = true;
                break;
            }
        }
        cur = (c == prev_c) ? prev_err : evaluate(c);
        prev_c = c;
        prev_err = cur;
        if (cur == cur && !isinf(cur)) {
            const double cc = cur * 1000.0, nearest = round(cc);
            tie = tie || (fabs(cc - nearest) < 2e-4 && nearest == (double)E + 1.0);
        }
        tie_topk = tie_topk || ((cutmask >> (it - 1)) & 1u);
        if (it <= 17)
            jump += hstep;
        else if (it <= 22)
            jump += tstep;
        else
            break;
    }
    const uint32_t ns = nsmall(c);
    if (t == 0) {
        fw->fft_count = c;
        fw->fft_err = cur;
        fw->fft_size = fft_payload_size(c, ns);
        fw->fft_iters = (uint16_t)it;
        fw->fft_tie = (tie ? TIE_FFT_LOOP : 0) | (tie_topk ? TIE_FFT_TOPK : 0);
        fw->fft_valid = pruned ? 2 : 1;
    }
}

}  // namespace atsc
