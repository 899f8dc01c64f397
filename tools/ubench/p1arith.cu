// Microbenchmark: the per-sample arithmetic of k_poly1 / k_poly1s (Hermite value, round to 1e-5, clamp, MAPE term)
// from registers only -- what the SM sustains with no memory access at all, and which part costs what.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I atsc_b200/csrc -o p1arith tools/ubench/p1arith.cu && ./p1arith
#include <cstdio>
#include <cuda_runtime.h>
#include "poly.cuh"
using namespace atsc;
// VAR 0: full tame term; 1: no reciprocal (MUFU + 2 DFMA -> 1 DMUL); 2: no FRND (round step skipped);
// 3: no clamp (2 DSETP + 4 FSEL); 4: Hermite only; 5: full non-tame term
template <int VAR>
__global__ void __launch_bounds__(512, 2) k(double *out, long long *cyc, double seed, int iters) {
    const double h00 = seed * 0.3 + threadIdx.x * 1e-4, h10 = seed * 0.1, h01 = 1.0 - h00, h11 = -seed * 0.05;
    const double vmin = seed * 0.5, vmax = seed * 40.0;
    double kv[5], tv[5], o[4];
#pragma unroll
    for (int i = 0; i < 5; i++) kv[i] = seed * (10.0 + i) + threadIdx.x * 1e-3, tv[i] = seed * 0.01 * (i + 1);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = seed * (10.5 + i);
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(kv[u], h00), __dmul_rn(tv[u], h10)), __dmul_rn(kv[u + 1], h01)),
                                       __dmul_rn(tv[u + 1], h11));
            double r;
            if (VAR == 4) { e[u] = v; continue; }
            if (VAR == 5) { e[u] = mape_term(round_and_limit5_fast(v, vmin, vmax), o[u]); continue; }
            if (VAR == 2) r = div_1e5_int53(__dmul_rn(v, 100000.0));
            else r = div_1e5_int53(round_half_away(__dmul_rn(v, 100000.0)));
            if (VAR != 3) { if (r < vmin) r = vmin; else if (r > vmax) r = vmax; }
            if (VAR == 1) e[u] = fabs(__dmul_rn(__dsub_rn(r, o[u]), o[u]));
            else e[u] = mape_term_tame(r, o[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += e[u];
        // new operands every trip: one 32-bit integer add on the low mantissa word of each (ALU pipe, not FP64)
#pragma unroll
        for (int i = 0; i < 5; i++) {
            kv[i] = __hiloint2double(__double2hiint(kv[i]), __double2loint(kv[i]) + 0x1357 * (i + 1));
            tv[i] = __hiloint2double(__double2hiint(tv[i]), __double2loint(tv[i]) + 0x2468 * (i + 1));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = __hiloint2double(__double2hiint(o[i]), __double2loint(o[i]) + 0x369c * (i + 1));
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VAR>
void run(const char *name, int ctas_per_sm) {
    double *out; long long *cyc;
    const int nb = 148 * ctas_per_sm;
    cudaMalloc(&out, nb * 512 * 8); cudaMalloc(&cyc, nb * 8);
    int iters = 4000;
    k<VAR><<<nb, 512>>>(out, cyc, 1.25, iters);
    k<VAR><<<nb, 512>>>(out, cyc, 1.25, iters);
    long long h[296];
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double samples = 512.0 * ctas_per_sm * iters * 4;
    long long mx = 0; for (int i = 0; i < nb; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s %d x 512 threads/SM: %6.3f samples/cycle/SM (%lld cycles)\n", name, ctas_per_sm, samples / mx, mx);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int c : {1, 2}) {
        run<0>("full tame term", c);
        run<1>("no reciprocal (MUFU + 2 DFMA -> DMUL)", c);
        run<2>("no FRND", c);
        run<3>("no clamp", c);
        run<4>("Hermite value only", c);
        run<5>("full non-tame term", c);
    }
    return 0;
}
