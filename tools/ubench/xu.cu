// Microbenchmark: per-SM throughput (thread-ops per cycle) of the FP64-related instructions k_front / k_poly lean on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu xu.cu && ./xu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, long long *cyc, double seed, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-3 + i * 0.37;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = trunc(a[i] * 1.0000001);                 // DMUL + FRND.F64.TRUNC
            if (OP == 1) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y + 1.5; }  // MUFU.RCP64H + DADD
            if (OP == 2) a[i] = (double)(float)a[i] * 1.0000001;            // F2F.F32.F64 + F2F.F64.F32 + DMUL
            if (OP == 3) a[i] = __fma_rn(a[i], 1.0000001, 1e-9);            // DFMA
            if (OP == 4) a[i] = __dmul_rn(a[i], 1.0000001);                 // DMUL (baseline for 0)
            if (OP == 5) a[i] = __dadd_rn(a[i], 1.5);                       // DADD (baseline for 1)
            if (OP == 6) a[i] = (a[i] > 3.0) ? a[i] * 0.5 : a[i] * 1.7;     // DSETP + DMUL + select
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char *name, int threads) {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    int iters = 2000;
    k<OP><<<148, threads>>>(out, cyc, 1.25, iters);
    k<OP><<<148, threads>>>(out, cyc, 1.25, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double ops = (double)threads * iters * 8;
    printf("%-34s threads/SM %4d: %8.2f thread-ops/cycle/SM (%lld cycles)\n", name, threads, ops / h[0], h[0]);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {512, 1024}) {
        run<0>("DMUL + FRND.F64.TRUNC", th);
        run<4>("DMUL", th);
        run<1>("MUFU.RCP64H + DADD", th);
        run<5>("DADD", th);
        run<2>("F2F.F32.F64 + F2F.F64.F32 + DMUL", th);
        run<3>("DFMA", th);
        run<6>("DSETP + 2 DMUL + sel", th);
    }
    return 0;
}
