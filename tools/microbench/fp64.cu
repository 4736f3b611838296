// FP64 latency / throughput probe for the B200 used by this pool (informs the Swing fit design).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_chain(double *out, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = __fma_rn(x, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / iters; }
}
__global__ void indep8(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7; out[1] = (double)(t1 - t0) / (8.0 * iters); }
}
__global__ void div_chain(double *out, int iters, double a) {
    double x = 1e300 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = __ddiv_rn(x, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / iters; }
}
__global__ void f32_chain(double *out, int iters, float a, float b) {
    float x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = __fmaf_rn(x, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / iters; }
}
int main() {
    double *d, h[2];
    cudaMalloc(&d, 16);
    auto rd = [&](const char *name) { cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-44s %8.2f cycles/op\n", name, h[1]); };
    dep_chain<<<1, 32>>>(d, 100000, 1.0000001, 1e-9); rd("DFMA dependent chain, 1 warp");
    indep8<<<1, 32>>>(d, 100000, 1.0000001, 1e-9); rd("DFMA 8 independent, 1 warp (per DFMA)");
    indep8<<<1, 128>>>(d, 100000, 1.0000001, 1e-9); rd("DFMA 8 independent, 4 warps/SM (per DFMA/warp)");
    indep8<<<1, 1024>>>(d, 100000, 1.0000001, 1e-9); rd("DFMA 8 independent, 32 warps/SM (per DFMA/warp)");
    div_chain<<<1, 32>>>(d, 20000, 1.0000001); rd("DDIV dependent chain, 1 warp");
    f32_chain<<<1, 32>>>(d, 100000, 1.0000001f, 1e-9f); rd("FFMA dependent chain, 1 warp");
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    indep8<<<148 * 4, 1024>>>(d, 1000, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); indep8<<<148 * 4, 1024>>>(d, 20000, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * 20000.0 * 148 * 4 * 1024;
    printf("full-chip DFMA throughput: %.2f TFLOP/s\n", flops / (ms * 1e-3) / 1e12);
    return 0;
}
