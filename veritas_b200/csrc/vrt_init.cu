// Rectangle::InitializeDistribution (Rectangle.cpp:616-665) on the device for the shipped Maxwellian slab
// (Settings::InitialDistribution, veritas.cpp:107-115): sub-cell midpoint quadrature, nvals = r^(depth+quadratureDepth)
// points per direction.  SURVEY.md §8(f) item 2 ("next"): lets large meshes start without a host pass.  The quadrature points
// are formed in the reference's association — ((0.5 + k)/nvals + i + x_pos)·dx and Momentum((0.5 + l)/nvals + j) =
// pmin + dp·(((0.5 + l)/nvals + j) + p_pos) (Rectangle.cpp:652-653, Rectangle.hpp:86-88) — and this unit is compiled with
// -fmad=false, so the only difference to a host-initialised run is device exp() vs libm's (<= 1 ulp);
// tests/test_gpu_initial_condition.py compares with the reference's step-0 state.
#include "vrt_internal.cuh"
#include <cmath>

namespace {
#define VRT_DPI 6.28318530718   // veritas.hpp:29

__device__ double cell_average(int i, int x_pos, double dx, double pmin, double dp, int j, int p_pos, int nvals, double xl, double xr, double n0, double T) {
    double temp = 0.0;
    const double norm = sqrt(VRT_DPI * T);
    for (int k = 0; k < nvals; k++) {
        double xp = ((0.5 + k) / nvals + i + x_pos) * dx;
        double ne = ((xp > xl) && (xp < xr)) ? n0 : 0.0;
        for (int l = 0; l < nvals; l++) {
            double pp = pmin + dp * (((0.5 + l) / nvals + j) + p_pos);
            temp += ne * exp(-(pp * pp) / (2.0 * T)) / norm;
        }
    }
    return temp * (1 / ((double)nvals * (double)nvals));
}

__global__ void k_init_patch(VrtPatchDev P, double pmin, int nvals, double xl, double xr, double n0, double T) {
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.npad) return;
    int i = (int)(c / P.pitch) - 2, j = (int)(c % P.pitch) - 2;
    if (i < 0 || i >= P.n_x || j < 0 || j >= P.n_p) return;
    double v = cell_average(i, P.x_pos, P.dx, pmin, P.dp, j, P.p_pos, nvals, xl, xr, n0, T);
    P.f0[c] = v; P.f1[c] = v;
}

__global__ void k_init_slab(VrtSlabDev L, double* plane, double pmin, int nvals, double xl, double xr, double n0, double T) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long ncol = L.n_x + 2 * L.gx;
    if (idx >= ncol * L.n_p) return;
    int cl = (int)(idx / L.n_p) - L.gx, j = (int)(idx % L.n_p);
    int gi = L.x_begin + cl;
    if (gi < 0 || gi >= L.n_x_global) return;   // physical ghost columns stay 0
    plane[(long)(cl + L.gx) * L.pitch + VRT_SLAB_GH + j] = cell_average(gi, 0, L.dx, pmin, L.dp, j, 0, nvals, xl, xr, n0, T);
}
}  // namespace

int vrt_init_kernels_maxwellian(vrt_ctx* c, int s, double xl, double xr, double n0, double T, int quadrature_depth) {
    VrtSpeciesState& S = c->S[s];
    const int r = c->refinement_ratio;
    if (S.path == VRT_PATH_FUSED) {
        VrtSlabDev& L = S.slab;
        int nvals = (int)std::lround(std::pow((double)r, quadrature_depth));
        S.i_f1 = S.i_f0;
        long n = (long)(L.n_x + 2 * L.gx) * L.n_p;
        k_init_slab<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(L, L.f[S.i_f0], S.sp.pmin, nvals, xl, xr, n0, T);
        c->launches += 1;
        VRT_CUDA(c, cudaGetLastError());
        return 0;
    }
    for (const VrtPatchDev& P : S.table) {
        int nvals = (int)std::lround(std::pow((double)r, P.depth + quadrature_depth));
        k_init_patch<<<(unsigned)((P.npad + 255) / 256), 256, 0, c->stream>>>(P, S.sp.pmin, nvals, xl, xr, n0, T);
        c->launches += 1;
        VRT_CUDA(c, cudaGetLastError());
    }
    return 0;
}
