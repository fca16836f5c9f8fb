// Split path: Rectangle::FCTTimeStep sub-steps as separate kernels over the patches of one level, with the
// reference's loop bounds, so that ghost/limiter syncs can run between them exactly as Mesh::Advance orders
// them (Mesh.cpp:64-89).  Works for any hierarchy; one launch per level and sub-step through a
// patch-descriptor table (blockIdx.y = patch).  Compiled with -fmad=false: bit-compatible with the
// reference's non-FMA build in everything except libm log and the p-reduction order of the moments.
#include "vrt_internal.cuh"
#include "vrt_launch.cuh"
#include "vrt_device.cuh"
#include <algorithm>

namespace {

__constant__ VrtTableau c_tabs;
bool g_tabs_loaded[64] = {};     // __constant__ memory is per device: one flag per device ordinal

#define PATCH_THREAD_SETUP                                                       \
    const VrtPatchDev& P = patches[blockIdx.y];                                  \
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;                        \
    if (c >= P.npad) return;                                                     \
    const int i = (int)(c / P.pitch) - 2, j = (int)(c % P.pitch) - 2;            \
    const int nx = P.n_x, np = P.n_p; (void)nx; (void)np; (void)i; (void)j;

// sub-step 0, part 1: advection speeds and WENO face values (Rectangle.cpp:1276-1312)
__global__ void k_speeds_faces(const VrtPatchDev* patches, Sp sp, VrtFields F) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    const double q = sp.q, dx_inv = 1 / P.dx, dp_inv = 1 / P.dp, cc = VRT_CS * VRT_CS * sp.m;
    if (i >= -1 && i <= nx + 1 && j >= -1 && j <= np) {
        double as = q * q * a_sq(F, finest_index(P, i));
        double am = __dmul_rn(__dmul_rn(dp_inv, cc), __dadd_rn(gamma_(sp, momentum(P, sp, j + 1), as), -gamma_(sp, momentum(P, sp, j), as)));
        P.ex[c] = am;
        if (i >= 0 && i <= nx)
            P.fx[c] = weno(P.f1[NS(P, i - 2, j)], P.f1[NS(P, i - 1, j)], P.f1[c], P.f1[NS(P, i + 1, j)], am > 0.0);
    }
    if (i >= -1 && i <= nx && j >= -1 && j <= np + 1) {
        double as_1 = q * q * a_sq(F, finest_index(P, i));
        double as_2 = q * q * a_sq(F, finest_index(P, i + 1));
        double Em = q * patch_efield(P, F, i);
        double mom = momentum(P, sp, j);
        double am = __dadd_rn(Em, -__dmul_rn(__dmul_rn(cc, dx_inv), __dadd_rn(gamma_(sp, mom, as_2), -gamma_(sp, mom, as_1))));
        P.ep[c] = am;
        if (j >= 0 && j <= np)
            P.fp[c] = weno(P.f1[c - 2], P.f1[c - 1], P.f1[c], P.f1[c + 1], am > 0.0);
    }
}

// sub-step 0, part 2: high/low-order fluxes at faces not flagged as interior level boundaries (Rectangle.cpp:1313-1394;
// flagged faces are filled by k_level_boundary_fluxes of vrt_amr.cu).  FxL/FpL hold slot 0 only (quirk Q1).
__global__ void k_fluxes(const VrtPatchDev* patches, int step) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    const double w3 = 1 / 48.0, dx_inv = 1 / P.dx, dp_inv = 1 / P.dp;
    const unsigned char fl = P.flags[c];
    double* FxHs = P.FxH + step * P.npad; double* FpHs = P.FpH + step * P.npad;
    if (i >= 0 && i <= nx && j >= -1 && j <= np && !(fl & VRT_LBX)) {
        double am = P.ex[c], ap1 = P.ex[c + 1], am1 = P.ex[c - 1];
        double fm = P.fx[c], fp1 = P.fx[c + 1], fm1 = P.fx[c - 1];
        FxHs[c] = dx_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
        if (step == 0) P.FxL[c] = dx_inv * ((am > 0.0 ? P.f1[NS(P, i - 1, j)] : P.f1[c]) * am);
    }
    if (i >= -1 && i <= nx && j >= 0 && j <= np && !(fl & VRT_LBP)) {
        long cp = c + P.pitch, cm = c - P.pitch;
        double am = P.ep[c], ap1 = P.ep[cp], am1 = P.ep[cm];
        double fm = P.fp[c], fp1 = P.fp[cp], fm1 = P.fp[cm];
        FpHs[c] = dp_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
        if (step == 0) P.FpL[c] = dp_inv * ((am > 0.0 ? P.f1[c - 1] : P.f1[c]) * am);
    }
}
// sub-step 0, parts 3 + 4 in one launch: the RK combination of every padded face (Rectangle.cpp:1396-1517) and, for interior cells, the
// low-order predictor f2 = f0 + FLS in gather form (Rectangle.cpp:1518-1534; quirks Q11, Q14: the reference's summation order).  The predictor needs FLS at the cell's own and its upper neighbours' faces; FLS = (sum_k a_sk) dt FL
// is one multiplication of a plane that is final before this kernel starts, so the neighbours' values are formed here again (the
// same product, the same bits) instead of being read back after a grid-wide dependency — one launch less on the chain.
__global__ void k_rk_combine_apply(const VrtPatchDev* patches, int step, const double* d_dt) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    const double timestep = *d_dt;
    double a[6], aSum = 0.0;
    for (int k = 0; k <= step; k++) { a[k] = c_tabs.a[step][k] * timestep; aSum = (k == 0) ? a[0] : aSum + a[k]; }
    double xl = aSum * P.FxL[c], pl = aSum * P.FpL[c];
    double sx = a[0] * P.FxH[c], sp_ = a[0] * P.FpH[c];
    for (int k = 1; k <= step; k++) { sx = sx + a[k] * P.FxH[k * P.npad + c]; sp_ = sp_ + a[k] * P.FpH[k * P.npad + c]; }
    P.FxLS[c] = xl; P.FpLS[c] = pl;
    P.FxDS[c] = sx - xl; P.FpDS[c] = sp_ - pl;
    if (i < 0 || i >= nx || j < 0 || j >= np) return;
    const int xm = P.left ? 1 : 0, xp = P.right ? nx : nx + 1, pp = P.up ? np : np + 1, pm = P.down ? 1 : 0;
    const long cxp = c + P.pitch;
    const bool in_i = (i >= xm && i < xp), in_j = (j >= pm && j < pp);
    const bool in_j1 = (j + 1 >= pm && j + 1 < pp), in_i1 = (i + 1 >= xm && i + 1 < xp);
    double v = P.f0[c];
    if (in_i && in_j) { v += xl; v += pl; }
    if (in_i && in_j1) v -= aSum * P.FpL[c + 1];
    if (in_i1 && in_j) v -= aSum * P.FxL[cxp];
    P.f2[c] = v;
}

// sub-step 2: f1 = f2 + C FDS in gather form, same summation order as the reference's serial scatter (Rectangle.cpp:1595-1612;
// quirks Q11, Q14)
__global__ void k_apply(const VrtPatchDev* patches) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    if (i < 0 || i >= nx || j < 0 || j >= np) return;
    const int xm = P.left ? 1 : 0, xp = P.right ? nx : nx + 1, pp = P.up ? np : np + 1, pm = P.down ? 1 : 0;
    const long cxp = c + P.pitch;
    const bool in_i = (i >= xm && i < xp), in_j = (j >= pm && j < pp);
    const bool in_j1 = (j + 1 >= pm && j + 1 < pp), in_i1 = (i + 1 >= xm && i + 1 < xp);
    double v = P.f2[c];
    if (in_i && in_j) { v += P.Cx[c] * P.FxDS[c]; v += P.Cp[c] * P.FpDS[c]; }
    if (in_i && in_j1) v -= P.Cp[c + 1] * P.FpDS[c + 1];
    if (in_i1 && in_j) v -= P.Cx[cxp] * P.FxDS[cxp];
    P.f1[c] = v;
}

// sub-step 1: Zalesak ratios R+- on [-1,n_x]x[-1,n_p] (Rectangle.cpp:1536-1579)
__global__ void k_limiter_r(const VrtPatchDev* patches) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    if (i < -1 || i > nx || j < -1 || j > np) return;
    const long cxp = c + P.pitch, cxm = c - P.pitch;
    double Pp = vmax(0.0, P.FxDS[c]) - vmin(0.0, P.FxDS[cxp]) + vmax(0.0, P.FpDS[c]) - vmin(0.0, P.FpDS[c + 1]);
    double Pm = vmax(0.0, P.FxDS[cxp]) - vmin(0.0, P.FxDS[c]) + vmax(0.0, P.FpDS[c + 1]) - vmin(0.0, P.FpDS[c]);
    double w1a = vmax(P.f0[c], P.f2[c]), w2a = vmax(P.f0[cxp], P.f2[cxp]), w3a = vmax(P.f0[cxm], P.f2[cxm]);
    double w4a = vmax(P.f0[c + 1], P.f2[c + 1]), w5a = vmax(P.f0[c - 1], P.f2[c - 1]);
    double wMax = vmax(w1a, vmax(w2a, vmax(w3a, vmax(w4a, w5a))));
    double w1i = vmin(P.f0[c], P.f2[c]), w2i = vmin(P.f0[cxp], P.f2[cxp]), w3i = vmin(P.f0[cxm], P.f2[cxm]);
    double w4i = vmin(P.f0[c + 1], P.f2[c + 1]), w5i = vmin(P.f0[c - 1], P.f2[c - 1]);
    double wMin = vmin(w1i, vmin(w2i, vmin(w3i, vmin(w4i, w5i))));
    double Qm = -wMin + P.f2[c], Qp = wMax - P.f2[c];
    P.Rp[c] = Pp > 0.0 ? vmin(1.0, Qp / Pp) : 0.0;
    P.Rm[c] = Pm > 0.0 ? vmin(1.0, Qm / Pm) : 0.0;
}
// sub-step 1: limiter C, 1.0 everywhere then the interior loop i in [1,n_x), j in [0,n_p) (Rectangle.cpp:1581-1594, quirk Q13)
__global__ void k_limiter_c(const VrtPatchDev* patches) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    double cx = 1.0, cp = 1.0;
    if (i >= 1 && i < nx && j >= 0 && j < np) {
        long ci = c - P.pitch, cj = c - 1;
        cx = P.FxDS[c] > 0.0 ? vmin(P.Rp[c], P.Rm[ci]) : vmin(P.Rp[ci], P.Rm[c]);
        cp = P.FpDS[c] > 0.0 ? vmin(P.Rp[c], P.Rm[cj]) : vmin(P.Rp[cj], P.Rm[c]);
    }
    P.Cx[c] = cx; P.Cp[c] = cp;
}
// sub-step 3: f0 := f1 over the padded array (Rectangle.cpp:1614-1622)
__global__ void k_commit(const VrtPatchDev* patches) { vrt_pdl_sync();
    PATCH_THREAD_SETUP
    P.f0[c] = P.f1[c];
}
// ---- moments: Rectangle::CalculateRhoAndJ, USINGMKL branch (Rectangle.cpp:157-282) -------------------------
__constant__ double c_IM[12] = {0.104166666666667, -0.708333333333334, 0.708333333333334, -0.104166666666667,
                                0.117647058823529, 0.029411764705882,  0.029411764705882, 0.117647058823529,
                                -0.083333333333333, 0.166666666666667, -0.166666666666667, 0.083333333333334};
// sub-cell coefficients of GetInterpolantsREL (Rectangle.cpp:139-155) for sub-cell kk of rtb: they depend on (kk, rtb) only, so a
// block tabulates them once in shared memory (same expressions, same bits) instead of re-deriving them — two fp64 divisions and
// ~15 operations — for every sub-cell of every one of the three interpolations a cell needs
constexpr int RTB_TAB = 32;
__device__ __forceinline__ void rel_coefficients(int kk, int rtb, double& c0, double& c1, double& c2) {
    double tl = -0.5 + kk / (double)rtb, tr = -0.5 + (kk + 1.0) / (double)rtb;
    c0 = (tl + tr) * 0.5;
    c1 = (tl * tl + tl * tr + tr * tr) / 3.0 - (1.0 / 12);
    c2 = (tl * tl * tl + tl * tl * tr + tl * tr * tr + tr * tr * tr) * 0.25;
}
// one sub-cell k of GetInterpolantsREL plus the mean-preserving correction (Rectangle.cpp:198-205); tab = [3][RTB_TAB] or nullptr
__device__ double rel_value(const VrtPatchDev& P, int i, int j, int k, const double* tab) {
    const int rtb = P.rtb;
    double f1 = P.f1[NS(P, i - 2, j)], f2 = P.f1[NS(P, i - 1, j)], f3 = P.f1[NS(P, i, j)], f4 = P.f1[NS(P, i + 1, j)], f5 = P.f1[NS(P, i + 2, j)];
    const double fc = f3;
    f5 -= f3; f4 -= f3; f2 -= f3; f1 -= f3;
    double a1 = c_IM[0] * f1 + c_IM[1] * f2 + c_IM[2] * f4 + c_IM[3] * f5;
    double a2 = c_IM[4] * f1 + c_IM[5] * f2 + c_IM[6] * f4 + c_IM[7] * f5;
    double a3 = c_IM[8] * f1 + c_IM[9] * f2 + c_IM[10] * f4 + c_IM[11] * f5;
    double sum = 0.0, mine = 0.0;
    for (int kk = 0; kk < rtb; kk++) {
        double c0, c1, c2;
        if (tab) { c0 = tab[kk]; c1 = tab[RTB_TAB + kk]; c2 = tab[2 * RTB_TAB + kk]; }
        else rel_coefficients(kk, rtb, c0, c1, c2);
        double v = c0 * a1 + c1 * a2 + c2 * a3 + f3;
        sum += v;
        if (kk == k) mine = v;
    }
    double cor = fc - (1.0 / (double)rtb) * sum;
    return mine + cor;
}
__device__ __forceinline__ double cell_a_sq(const VrtFields& F, int i) {   // EMSolver.hpp:56-63
    i += F.pre; i = i > -1 ? i : 0; i = i < F.M ? i : F.M - 1;
    double ay = F.Y[VRT_AY][F.M + i], az = F.Y[VRT_AZ][F.M + i];
    return (ay * ay) + (az * az);
}
// grid: (n_x*rtb, patches); block reduces over p.  A cell's sums need u = gamma + p/mc at its four surrounding faces and
// g = mc ln(u_hi/u_lo) of itself and its two p-neighbours: every thread forms u at ONE face and g of ONE cell and the block shares
// them through shared memory (the same expressions on the same operands as forming them per cell — four square roots, three
// divisions and three logarithms per cell before — so the values are the same bits).
__global__ void __launch_bounds__(128) k_moments(const VrtPatchDev* patches, Sp sp, VrtFields F) { vrt_pdl_sync();
    const VrtPatchDev& P = patches[blockIdx.y];
    const int rtb = P.rtb;
    if ((int)blockIdx.x >= P.n_x * rtb) return;
    const int i = blockIdx.x / rtb, k = blockIdx.x % rtb;
    const double q = sp.q, c1 = sp.m_inv * VRT_C_INV, c2 = 1 / c1, c3 = 1 / 48.0;
    const double a2 = q * q * cell_a_sq(F, (i + P.x_pos) * rtb + k);
    __shared__ double coef[3 * RTB_TAB];
    __shared__ double su[128 + 3], sg[128 + 2];       // u at faces j0 - 1 .. j0 + 129, g of cells j0 - 1 .. j0 + 128
    const double* tab = rtb <= RTB_TAB ? coef : nullptr;
    const int t = threadIdx.x;
    if (tab && t < rtb) rel_coefficients(t, rtb, coef[t], coef[RTB_TAB + t], coef[2 * RTB_TAB + t]);
    auto uface = [&](int j) { const double p = momentum(P, sp, j); return gamma_(sp, p, a2) + c1 * p; };
    double rho = 0.0, cur = 0.0;
    for (int j0 = 0; j0 < P.n_p; j0 += 128) {
        __syncthreads();                                // the previous tile's su / sg are no longer read (first pass: coef is written)
        su[t + 1] = uface(j0 + t);
        if (t < 3) su[t == 0 ? 0 : 128 + t] = uface(t == 0 ? j0 - 1 : j0 + 127 + t);      // faces j0 - 1, j0 + 128, j0 + 129
        __syncthreads();
        sg[t + 1] = c2 * log(su[t + 2] / su[t + 1]);                                       // cell j0 + t: faces j0 + t, j0 + t + 1
        if (t < 2) { const int e = t == 0 ? 0 : 129; sg[e] = c2 * log(su[e + 1] / su[e]); }   // cells j0 - 1 and j0 + 128
        __syncthreads();
        const int j = j0 + t;
        if (j < P.n_p && !(P.flags[NS(P, i, j)] & VRT_NESTED)) {   // cells covered by a finer patch are skipped (Rectangle.cpp:207-208)
            const double t0 = rel_value(P, i, j, k, tab), tm1 = rel_value(P, i, j - 1, k, tab), tp1 = rel_value(P, i, j + 1, k, tab);
            rho += t0;
            cur += t0 * sg[t + 1] + c3 * (sg[t + 2] - sg[t]) * (tp1 - tm1);
        }
    }
    __shared__ double sh[2][4];
    for (int o = 16; o > 0; o >>= 1) { rho += __shfl_down_sync(0xffffffffu, rho, o); cur += __shfl_down_sync(0xffffffffu, cur, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = rho; sh[1][threadIdx.x >> 5] = cur; }
    __syncthreads();
    if (threadIdx.x == 0) {
        rho = (sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3]);
        cur = (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]);
        P.chargeR[blockIdx.x] = rho * (P.dp * q);
        P.currentR[blockIdx.x] = cur * (-q * q / sp.m);
    }
}

// Rectangle::CalculateEnergy (Rectangle.cpp:284-305), race-free: energyR[j * rtb + k] = dx * sum over the non-nested cells of
// the patch's row j of sub-cell k of GetInterpolantsREL along p (no mean correction there).  One block per
// (p sub-cell, patch); threads stride over x, fixed-order tree.  The reference accumulates this inside `omp parallel for` over i
// without a reduction clause (a data race); the sum here is deterministic.
__global__ void __launch_bounds__(128) k_energy(const VrtPatchDev* patches, int patch, double* out) {
    const VrtPatchDev& P = patches[patch];
    const int rtb = P.rtb;
    const int j = blockIdx.x / rtb, k = blockIdx.x % rtb;
    const double tl = -0.5 + k / (double)rtb, tr = -0.5 + (k + 1.0) / (double)rtb;
    const double c0 = (tl + tr) * 0.5;
    const double c1 = (tl * tl + tl * tr + tr * tr) / 3.0 - (1.0 / 12);
    const double c2 = (tl * tl * tl + tl * tl * tr + tl * tr * tr + tr * tr * tr) * 0.25;
    double acc = 0.0;
    for (int i = threadIdx.x; i < P.n_x; i += blockDim.x) {
        const long c = NS(P, i, j);
        if (P.flags[c] & VRT_NESTED) continue;
        double f1 = P.f1[c - 2], f2 = P.f1[c - 1], f3 = P.f1[c], f4 = P.f1[c + 1], f5 = P.f1[c + 2];
        f5 -= f3; f4 -= f3; f2 -= f3; f1 -= f3;
        const double a1 = c_IM[0] * f1 + c_IM[1] * f2 + c_IM[2] * f4 + c_IM[3] * f5;
        const double a2 = c_IM[4] * f1 + c_IM[5] * f2 + c_IM[6] * f4 + c_IM[7] * f5;
        const double a3 = c_IM[8] * f1 + c_IM[9] * f2 + c_IM[10] * f4 + c_IM[11] * f5;
        acc += c0 * a1 + c1 * a2 + c2 * a3 + f3;
    }
    __shared__ double sh[4];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) * P.dx;
}
// the same on slab storage (rtb = 1: the interpolation is the identity): column sums over x in two passes
__global__ void k_slab_energy_partial(const double* f1p, int n_x, int n_p, int gx, int pitch, int chunk, double* partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_p) return;
    const int x0 = blockIdx.y * chunk, x1 = min(n_x, x0 + chunk);
    double acc = 0.0;
    for (int i = x0; i < x1; i++) acc += f1p[(long)(i + gx) * pitch + VRT_SLAB_GH + j];
    partial[(long)blockIdx.y * n_p + j] = acc;
}
__global__ void k_slab_energy_finish(const double* partial, int n_p, int n_chunks, double dx, double* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_p) return;
    double acc = 0.0;
    for (int k = 0; k < n_chunks; k++) acc += partial[(long)k * n_p + j];
    out[j] = acc * dx;
}

inline Sp make_sp(const VrtSpecies& s) { return Sp{s.m, s.q, s.pmin, 1 / s.m}; }

}  // namespace

// Rectangle::CalculateEnergy for one patch; host_energy receives n_p * rtb values (Rectangle::energyR)
int vrt_split_patch_energy(vrt_ctx* c, int s, int patch, double* host_energy) {
    VrtSpeciesState& S = c->S[s];
    double* d = nullptr;
    long n;
    if (S.path == VRT_PATH_FUSED) {
        const VrtSlabDev& L = S.slab;
        n = L.n_p;
        const int chunks = std::max(1, std::min(256, L.n_x / 64)), chunk = (L.n_x + chunks - 1) / chunks;
        VRT_CUDA(c, cudaMalloc(&d, sizeof(double) * n * (chunks + 1)));
        k_slab_energy_partial<<<dim3((unsigned)((n + 127) / 128), chunks), 128, 0, c->stream>>>(L.f[S.i_f1], L.n_x, L.n_p, L.gx, L.pitch, chunk, d + n);
        k_slab_energy_finish<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d + n, L.n_p, chunks, L.dx, d);
    } else {
        const VrtPatchDev& P = S.patches[patch];
        n = (long)P.n_p * P.rtb;
        VRT_CUDA(c, cudaMalloc(&d, sizeof(double) * n));
        k_energy<<<(unsigned)n, 128, 0, c->stream>>>(S.d_patches, S.table_index[patch], d);
    }
    c->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_energy, d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) { c->err = std::string("vrt_patch_energy: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
    return 0;
}

int vrt_split_init_tables(vrt_ctx* c) {
    bool& loaded = g_tabs_loaded[c->device & 63];
    if (!loaded) { VRT_CUDA(c, cudaMemcpyToSymbol(c_tabs, &kTableau, sizeof(VrtTableau))); loaded = true; }
    return 0;
}

static bool level_grid(const VrtSpeciesState& S, int depth, dim3* grid, int* first) {
    if (depth < 0 || depth >= (int)S.level_patches.size() || S.level_patches[depth].empty()) return false;
    long mx = 0;
    for (int p : S.level_patches[depth]) mx = std::max(mx, S.table[p].npad);
    *first = S.level_patches[depth][0];   // patches of one level are contiguous in the table
    *grid = dim3((unsigned)((mx + 255) / 256), (unsigned)S.level_patches[depth].size());
    return true;
}

int vrt_split_substep(vrt_ctx* c, int s, int depth, const double* d_dt, int step, int substep) {
    VrtSpeciesState& S = c->S[s];
    dim3 grid; int first;
    if (!level_grid(S, depth, &grid, &first)) return 0;
    const VrtPatchDev* tab = S.d_patches + first;
    Sp sp = make_sp(S.sp);
    if (substep == 0) {
        vrt_launch(k_speeds_faces, dim3(grid), dim3(256), c->stream, tab, sp, c->F);
        vrt_launch(k_fluxes, dim3(grid), dim3(256), c->stream, tab, step);
        c->launches += 2;
        if (int r = vrt_amr_level_boundary_fluxes(c, s, depth, step)) return r;
        vrt_launch(k_rk_combine_apply, dim3(grid), dim3(256), c->stream, tab, step, d_dt);
        c->launches += 1;
    } else if (substep == 1) {
        vrt_launch(k_limiter_r, dim3(grid), dim3(256), c->stream, tab);
        vrt_launch(k_limiter_c, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 2;
    } else if (substep == 2) {
        vrt_launch(k_apply, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 1;
    } else if (substep == 3) {
        vrt_launch(k_commit, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 1;
    } else {
        c->err = "vrt_vlasov_substep: substep must be 0..3";
        return VRT_ERR_ARG;
    }
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// The same sub-step for every level of the species in one launch (blockIdx.y runs over the whole patch table).  Inside a
// sub-step the levels do not interact: each kernel writes its own patch's planes and reads its own f, the 1-D fields and — in
// the coarse-fine flux matching — only f^(s) of finer patches, which no sub-step-0 kernel writes; the reference's
// coarse-to-fine level loop (Mesh.cpp:66-70) is therefore order-free.  Used by the stage sequence of vrt_step (fewer, fatter
// launches: small hierarchies are launch-latency bound); Level::FCTTimeStep keeps the per-level entry above.
int vrt_amr_level_boundary_fluxes_all(vrt_ctx* c, int s, int step);
int vrt_split_substep_all(vrt_ctx* c, int s, const double* d_dt, int step, int substep) {
    VrtSpeciesState& S = c->S[s];
    if (S.table.empty()) return 0;
    long mx = 0;
    for (const VrtPatchDev& T : S.table) mx = std::max(mx, T.npad);
    const dim3 grid((unsigned)((mx + 255) / 256), (unsigned)S.table.size());
    const VrtPatchDev* tab = S.d_patches;
    Sp sp = make_sp(S.sp);
    if (substep == 0) {
        vrt_launch(k_speeds_faces, dim3(grid), dim3(256), c->stream, tab, sp, c->F);
        vrt_launch(k_fluxes, dim3(grid), dim3(256), c->stream, tab, step);
        c->launches += 2;
        if (int r = vrt_amr_level_boundary_fluxes_all(c, s, step)) return r;
        vrt_launch(k_rk_combine_apply, dim3(grid), dim3(256), c->stream, tab, step, d_dt);
        c->launches += 1;
    } else if (substep == 1) {
        vrt_launch(k_limiter_r, dim3(grid), dim3(256), c->stream, tab);
        vrt_launch(k_limiter_c, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 2;
    } else if (substep == 2) {
        vrt_launch(k_apply, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 1;
    } else {
        vrt_launch(k_commit, dim3(grid), dim3(256), c->stream, tab);
        c->launches += 1;
    }
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// the species' patch moments (Rectangle::chargeR / currentR of every patch); the level sums and the species / total assembly are
// k_assemble's (vrt_fields.cu)
int vrt_split_moments(vrt_ctx* c, int s) {
    VrtSpeciesState& S = c->S[s];
    if (S.table.empty()) return 0;
    Sp sp = make_sp(S.sp);
    int mx = 0;
    for (const VrtPatchDev& T : S.table) mx = std::max(mx, T.n_x * T.rtb);
    vrt_launch(k_moments, dim3(dim3(mx, (unsigned)S.table.size())), dim3(128), c->stream, S.d_patches, sp, c->F);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}
