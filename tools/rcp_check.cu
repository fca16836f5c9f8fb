// accuracy of the seeded reciprocal used by the fused kernels (vrt_device.cuh: rcp_scaled) against IEEE division
#include <cstdio>
#include <cmath>
#include "../veritas_b200/csrc/vrt_device.cuh"
__global__ void k(double* out, int n) {
    double worst = 0, worst_seed = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = ldexp(1.0 + (double)i / n, (i % 41) - 20) * ((i & 1) ? 1.0 : 1.0000001);
        const double exact = 1.0 / d;
        double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
        worst_seed = fmax(worst_seed, fabs(x - exact) / exact);
        worst = fmax(worst, fabs(rcp_scaled(d) - exact) / exact);
    }
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(worst));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(worst_seed));
}
int main() {
    double* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    k<<<592, 256>>>(d, 1 << 26);
    double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("{\"rcp_scaled_max_rel_err\": %.3e, \"seed_max_rel_err\": %.3e, \"ulp\": %.3e}\n", h[0], h[1], ldexp(1.0, -53));
    return 0;
}
