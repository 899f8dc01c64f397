// FP64 dependent-chain throughput vs independent chains per thread (16 warps per SM, one CTA per SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH, int OP>
__global__ void k(double *out, long long *cyc, double seed, int iters) {
    double a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = seed + threadIdx.x * 1e-3 + i * 0.37;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (OP == 0) a[i] = __fma_rn(a[i], 1.0000001, 1e-9);
                if (OP == 1) a[i] = trunc(a[i]) + 0.3;
                if (OP == 2) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NCH, int OP>
void run(const char *name, int threads) {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    int iters = 1000;
    k<NCH, OP><<<148, threads>>>(out, cyc, 1.25, iters);
    k<NCH, OP><<<148, threads>>>(out, cyc, 1.25, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double per_chain_op = (double)h[0] / (iters * 8.0);   // cycles per dependent op of one chain (all chains interleaved)
    printf("%-10s warps %2d chains/thread %d: %6.1f cycles per round of %d ops/thread -> %5.1f thread-ops/cycle/SM\n", name, threads / 32, NCH,
           per_chain_op, NCH, (double)threads * NCH / per_chain_op);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>("DFMA", 32); run<1, 0>("DFMA", 128); run<1, 0>("DFMA", 512); run<2, 0>("DFMA", 512); run<4, 0>("DFMA", 512); run<8, 0>("DFMA", 512);
    run<1, 1>("FRND+DADD", 32); run<4, 1>("FRND+DADD", 512);
    run<1, 2>("RCP64H", 32); run<4, 2>("RCP64H", 512);
    return 0;
}
