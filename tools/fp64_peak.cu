// DFMA micro-benchmark: measured fp64 (non-tensor) peak of the device, the second roofline denominator
// (MEASURED_PEAKS.json has no fp64 figure).  8 independent FMA chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int r = 0; r < 6; r++) {
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3);
        if (r > 0 && fl > best) best = fl;
    }
    printf("{\"sms\": %d, \"fp64_fma_tflops\": %.3f, \"fp64_instr_per_s\": %.4e}\n", sms, best / 1e12, best / 2);
    return 0;
}
