// 1-D field solver on the device: EMFieldSolver of the reference (/root/reference/EMSolver.cpp).
// Compiled with -fmad=false: a_squared feeds Gamma(), whose rounding decides the upwind direction of the
// p~0 row, so every expression keeps the reference's operation order without contraction.
#include "vrt_internal.cuh"
#include "vrt_launch.cuh"
#include <cstdlib>

namespace {

__constant__ VrtTableau c_tab;

// ---- RGKCalculateRHS (EMSolver.cpp:479-553) ---------------------------------------------------------
__global__ void k_field_rhs(VrtFields F, int step, const VrtStepParams* prm) { vrt_pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = F.M, N = F.N, pre = F.pre, post = F.post;
    if (i >= M) return;
    const double *By = F.Y[VRT_BY] + M, *Bz = F.Y[VRT_BZ] + M, *Ey = F.Y[VRT_EY] + M, *Ez = F.Y[VRT_EZ] + M,
                 *Ay = F.Y[VRT_AY] + M, *Az = F.Y[VRT_AZ] + M;   // slot 1
    const long s2 = (long)(step + 2) * M + i;
    const double dx_inv = 1 / F.dx;
    double rBy = 0, rBz = 0, rEy = 0, rEz = 0, rAy = 0, rAz = 0;
    if (i == 0) {
        double by0 = prm->laser[2 * step], bz0 = prm->laser[2 * step + 1];
        rBy = dx_inv * (Ez[i + 1] - Ez[i]);
        rBz = -dx_inv * (Ey[i + 1] - Ey[i]);
        rEy = -dx_inv * (Bz[i] - bz0) * VRT_EPS0_INV * VRT_MU_INV;
        rEz = dx_inv * (By[i] - by0) * VRT_EPS0_INV * VRT_MU_INV;
        rAy = -Ey[i]; rAz = -Ez[i];
    } else if (i < pre || (i >= pre + N && i < pre + N + post - 2)) {
        rBy = dx_inv * (Ez[i + 1] - Ez[i]);
        rBz = -dx_inv * (Ey[i + 1] - Ey[i]);
        rEy = -dx_inv * (Bz[i] - Bz[i - 1]) * VRT_EPS0_INV * VRT_MU_INV;
        rEz = dx_inv * (By[i] - By[i - 1]) * VRT_EPS0_INV * VRT_MU_INV;
        rAy = -Ey[i]; rAz = -Ez[i];
    } else if (i >= pre && i < pre + N) {
        const double c1 = -1.0 / 24, c2 = 9.0 / 8.0;
        double Jv = F.J[i - pre];
        rBy = dx_inv * (c1 * (Ez[i + 2] - Ez[i - 1]) + c2 * (Ez[i + 1] - Ez[i]));
        rBz = -dx_inv * (c1 * (Ey[i + 2] - Ey[i - 1]) + c2 * (Ey[i + 1] - Ey[i]));
        rEy = -dx_inv * (c1 * (Bz[i + 1] - Bz[i - 2]) + c2 * (Bz[i] - Bz[i - 1])) * VRT_EPS0_INV * VRT_MU_INV - VRT_EPS0_INV * Jv * Ay[i];
        rEz = dx_inv * (c1 * (By[i + 1] - By[i - 2]) + c2 * (By[i] - By[i - 1])) * VRT_EPS0_INV * VRT_MU_INV - VRT_EPS0_INV * Jv * Az[i];
        rAy = -Ey[i]; rAz = -Ez[i];
    }   // last two cells stay 0
    F.Y[VRT_BY][s2] = rBy; F.Y[VRT_BZ][s2] = rBz; F.Y[VRT_EY][s2] = rEy;
    F.Y[VRT_EZ][s2] = rEz; F.Y[VRT_AY][s2] = rAy; F.Y[VRT_AZ][s2] = rAz;
}

// ---- RGKUpdateIntermediateSolution (EMSolver.cpp:204-338) --------------------------------------------
__global__ void k_field_update(VrtFields F, int step, const VrtStepParams* prm) { vrt_pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = F.M;
    if (i >= M) return;
    const double timestep = prm->dt;
    double* y = F.Y[blockIdx.y];
    if (step < 5) {
        double s = y[i] + (c_tab.a[step][0] * timestep) * y[2L * M + i];
        for (int k = 1; k <= step; k++) s = s + (c_tab.a[step][k] * timestep) * y[(long)(2 + k) * M + i];
        y[M + i] = s;
    } else {
        double s = y[M + i];
        for (int k = 0; k < 6; k++) {
            double ak = (k < 5) ? c_tab.a[4][k] * timestep : 0.0;
            double bk = c_tab.a[5][k] * timestep - ak;
            s = s + bk * y[(long)(2 + k) * M + i];
        }
        y[i] = s; y[M + i] = s;
    }
}

__device__ __forceinline__ double weno_unbiased(double f1, double f2, double f3, double f4) {   // EMSolver.cpp:565-585
    double fL = (1.0 / 6) * (-f1 + 5 * f2 + 2 * f3), fR = (1.0 / 6) * (2 * f2 + 5 * f3 - f4);
    double AL = f1 - 2 * f2 + f3, BL = f3 - f1, AR = f2 - 2 * f3 + f4, BR = f4 - f2;
    double bL = 4.0 / 3 * (AL * AL) + 0.5 * AL * BL + 0.25 * (BL * BL);
    double bR = 4.0 / 3 * (AR * AR) - 0.5 * AR * BR + 0.25 * (BR * BR);
    double mm = 1.0e-10;
    double oL = 0.5 / ((mm + bL) * (mm + bL)), oR = 0.5 / ((mm + bR) * (mm + bR));
    double wL = oL / (oL + oR), wR = oR / (oL + oR);
    return wL * fL + wR * fR;
}
// ---- InterpolateToFaces (EMSolver.cpp:555-619) -------------------------------------------------------
__global__ void k_field_faces(VrtFields F) { vrt_pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F.N) return;
    const double *Ay = F.Y[VRT_AY] + F.M + F.pre + i, *Az = F.Y[VRT_AZ] + F.M + F.pre + i;
    double ay = weno_unbiased(Ay[-2], Ay[-1], Ay[0], Ay[1]);
    double az = weno_unbiased(Az[-2], Az[-1], Az[0], Az[1]);
    F.a_squared[i] = (ay * ay) + (az * az);
}

// ---- UpdatePotential (EMSolver.cpp:156-192) ----------------------------------------------------------
// The reference solves  A x = b  with A = the dense matrix of EMSolver.cpp:28-67 as LAPACK (column-major) sees it:
// column 0 = e_0, column c>0 = the periodic 4th-order -d2 stencil (w = 1/12, -4/3, 5/2, -4/3, 1/12) centred on
// row c.  Hence x_0 = sum(b) and x_{r>0} = y_r where  L y = b - sum(b) e_0, y_0 = 0, L the periodic stencil
// operator.  L factors exactly as  T D  with  T = periodic tridiag(-1/12, 7/6, -1/12)  and  D = periodic
// tridiag(-1, 2, -1).  T^-1 is a convolution with C rho^|k| (rho = 7 - sqrt(48), C = 6/sqrt(48)); rho^18 < 3e-21,
// so 17 terms per side are exact to fp64.  D y = z with y_0 = 0 is two running sums.  O(N), no N x N matrix.
constexpr int PK = 17;

// The passes below are written for a "block" of PTHR threads working on tile `blk`.  WIDE = false: the block is the CTA (multi-CTA
// passes, the cluster kernel, the single-CTA kernel walking the tiles).  WIDE = true: the block is one of up to PSMALL groups of PTHR
// consecutive threads of a wide CTA, each on its own tile, synchronising among themselves with a named barrier and keeping their
// own copy of every shared work array — the same arithmetic per thread, warp and group, so the same bits.
constexpr int PTHR = 256, PEL = 4, PTILE = PTHR * PEL;
constexpr int PMAXT = 4096;      // tiles: N <= 4 Mi finest cells
constexpr int PSMALL = 4;        // tiles of a "short" grid (N <= 4096: BASELINE configs 1, 2, 4)
__host__ __device__ constexpr int psw(int e) { return e + (e >> 2); }                // padded position (see Blk::at)
__host__ __device__ constexpr int psw_len(int n) { return (psw(n) + 2) & ~1; }        // padded length of n entries, even
constexpr int PTILE_SW = psw_len(PTILE + 2 * 17);                                    // the convolution's staged tile (PK = 17), padded
template <bool WIDE> struct Blk {
    static constexpr int groups = WIDE ? PSMALL : 1;
    __device__ static __forceinline__ int tid() { return WIDE ? (int)(threadIdx.x & (PTHR - 1)) : (int)threadIdx.x; }
    __device__ static __forceinline__ int grp() { return WIDE ? (int)(threadIdx.x / PTHR) : 0; }
    __device__ static __forceinline__ void sync() {
        if (WIDE) asm volatile("bar.sync %0, %1;" ::"r"(grp() + 1), "n"(PTHR) : "memory");
        else __syncthreads();
    }
    // Position of entry i of an N-vector of the scratch.  A thread owns PEL = 4 consecutive entries, so in shared memory (WIDE) a
    // half-warp's 8-byte accesses to "entry k of my run" would be 32 bytes apart — four to a bank.  One pad entry after every four
    // (i + i/4: 40 bytes apart, 5 coprime to 16) spreads them over all banks.  Global scratch (WIDE = false) stays dense.
    __device__ static __forceinline__ int at(int i) { return WIDE ? i + (i >> 2) : i; }
    __device__ static __forceinline__ int stride(int N) { return WIDE ? psw_len(N) : N; }
};

template <bool WIDE>
__device__ double block_sum(double v, double* sh) {
    using B = Blk<WIDE>;
    const int tid = B::tid();
    B::sync();
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sh[tid >> 5] = v;
    B::sync();
    if (tid < 32) {
        double t = tid < (PTHR >> 5) ? sh[tid] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (tid == 0) sh[32] = t;
    }
    B::sync();
    return sh[32];
}
// exclusive prefix over the per-thread totals of a block; returns the offset for this thread
template <bool WIDE>
__device__ double block_excl_scan(double v, double* sh, double* total) {
    using B = Blk<WIDE>;
    B::sync();
    const int lane = B::tid() & 31, w = B::tid() >> 5;
    double inc = v;
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sh[w] = inc;
    B::sync();
    if (w == 0) {
        double t = lane < (PTHR >> 5) ? sh[lane] : 0.0, ti = t;
        for (int o = 1; o < 32; o <<= 1) { double u = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += u; }
        sh[lane] = ti - t;               // exclusive warp offsets
        if (lane == 31) sh[32] = ti;     // block total
    }
    B::sync();
    *total = sh[32];
    return sh[w] + (inc - v);
}

// Five small multi-CTA passes (tiles of PTILE entries, PTHR threads, PEL consecutive entries per thread; every cross-tile
// quantity is a fixed-order sum of per-tile partials that each CTA recomputes for itself, so there are no atomics and the
// result does not depend on scheduling — all ranks of a multi-GPU run get the same bits):
//   k_poisson_rhs   b = (rho + rho_neutral) dx^2/eps0, tile sums            (EMSolver.cpp:160-166)
//   k_poisson_conv  b' = b - sum(b) e_0,  z = T^-1 b', tile sums
//   k_poisson_scan1 c = inclusive prefix of z, tile sums
//   k_poisson_dsum  tile sums of d = mean(c) - c
//   k_poisson_scan2 PHI_0 = sum(b), PHI_i = sum_{k<i} d_k

// sum of part[0..G) (returned to all threads) and, for this block, of part[0..blk): thread t owns a contiguous run of partials
template <bool WIDE>
__device__ double partial_prefix(const double* part, int G, int blk, double* sh, double* total) {
    using B = Blk<WIDE>;
    __shared__ double own[B::groups];
    const int cp = (G + PTHR - 1) / PTHR, lo = B::tid() * cp, hi = min(G, lo + cp);
    double loc = 0.0, before = 0.0;
    for (int k = lo; k < hi; k++) { if (k == blk) before = loc; loc += part[k]; }
    if (WIDE) {
        // G <= PSMALL: every partial sits in the first warp of the group (one per thread), the other warps hold zeros.  The general
        // scan below then reduces to the first warp's shuffle scan plus additions of +0.0 — written out here with exactly those
        // additions (x + 0.0 is not an identity for -0.0 and is not folded away), without the second phase and two of the barriers.
        __shared__ double tot_[B::groups];
        B::sync();
        if (B::tid() < 32) {
            const int lane = B::tid();
            double inc = loc;
            for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            const double all = __shfl_sync(0xffffffffu, inc, 31);
            if (lane == blk) own[B::grp()] = (0.0 + (inc - loc)) + before;
            if (lane == 0) tot_[B::grp()] = 0.0 + all;
        }
        B::sync();
        *total = tot_[B::grp()];
        return own[B::grp()];
    }
    const double off = block_excl_scan<WIDE>(loc, sh, total);
    if (blk >= lo && blk < hi) own[B::grp()] = off + before;
    B::sync();
    return own[B::grp()];
}

template <bool WIDE>
__device__ __forceinline__ void d_poisson_rhs(VrtFields F, double* part, const int blk) {
    using B = Blk<WIDE>;
    __shared__ double sh_[B::groups][40];
    double* sh = sh_[B::grp()];
    const int N = F.N, i0 = blk * PTILE + B::tid() * PEL, p0 = B::at(i0);      // i0 is a multiple of 4: entry i0 + k sits at p0 + k
    const double te = VRT_EPS0_INV, w = F.dx * F.dx;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PEL; k++) {
        const int i = i0 + k;
        if (i < N) { double v = te * (F.charge[i] + F.neutral[i]); v *= w; F.scratch[p0 + k] = v; s += v; }
    }
    const double t = block_sum<WIDE>(s, sh);
    if (B::tid() == 0) part[blk] = t;
}

template <bool WIDE>
__device__ __forceinline__ void d_poisson_conv(VrtFields F, const double* part_b, double* part_z, int G, const int blk, double* tiles = nullptr) {
    using B = Blk<WIDE>;
    static_assert(PTILE_SW == psw_len(PTILE + 2 * PK), "PTILE_SW");
    __shared__ double sh_[B::groups][40];
    __shared__ double green_[B::groups][PK + 1];
    __shared__ double tile_[WIDE ? 2 : PTILE_SW];                  // WIDE: the groups' tiles live in the dynamic allocation (`tiles`)
    double* sh = sh_[B::grp()]; double* green = green_[B::grp()]; double* tile = WIDE ? tiles + B::grp() * PTILE_SW : tile_;
    const int N = F.N, tid = B::tid(), t0 = blk * PTILE;
    const double* b = F.scratch;
    double* z = F.scratch + B::stride(N);
    if (tid <= PK) {
        const double rho = 1.0 / (7.0 + sqrt(48.0)), Cg = 6.0 / sqrt(48.0);   // rho = 7 - sqrt(48) without cancellation
        double g = Cg;
        for (int k = 0; k < tid; k++) g *= rho;
        green[tid] = g;
    }
    double sb;
    partial_prefix<WIDE>(part_b, G, 0, sh, &sb);
    for (int e = tid; e < PTILE + 2 * PK; e += PTHR) {
        int i = t0 - PK + e;                       // periodic wrap (N may be smaller than the halo)
        while (i < 0) i += N;
        while (i >= N) i -= N;
        const double v = b[B::at(i)];
        tile[psw(e)] = (i == 0) ? v - sb : v;
    }
    B::sync();
    double s = 0.0;
    // tile entry tid * PEL + c sits at 5 tid + psw(c) (the padded layout above: no bank conflicts between the threads of a warp)
    const double* mine = tile + 5 * tid;
    const int zp0 = B::at(t0 + tid * PEL);
#pragma unroll
    for (int k = 0; k < PEL; k++) {
        const int i = t0 + tid * PEL + k;
        if (i < N) {
            double acc = green[0] * mine[psw(k + PK)];
#pragma unroll
            for (int d = 1; d <= PK; d++) acc += green[d] * (mine[psw(k + PK + d)] + mine[psw(k + PK - d)]);
            z[zp0 + k] = acc; s += acc;
        }
    }
    const double t = block_sum<WIDE>(s, sh);
    if (tid == 0) part_z[blk] = t;
    if (blk == 0 && tid == 0) F.scratch[3L * B::stride(N)] = sb;
}

template <bool WIDE>
__device__ __forceinline__ void d_poisson_scan1(VrtFields F, const double* part_z, double* part_c, int G, const int blk) {
    using B = Blk<WIDE>;
    __shared__ double sh_[B::groups][40];
    double* sh = sh_[B::grp()];
    const int N = F.N, tid = B::tid(), i0 = blk * PTILE + tid * PEL, p0 = B::at(i0);
    const double* z = F.scratch + B::stride(N);
    double* cpre = F.scratch + 2L * B::stride(N);
    double tot;
    const double base = partial_prefix<WIDE>(part_z, G, blk, sh, &tot);
    double v[PEL], loc = 0.0;
#pragma unroll
    for (int k = 0; k < PEL; k++) { v[k] = (i0 + k < N) ? z[p0 + k] : 0.0; loc += v[k]; }
    double run = base + block_excl_scan<WIDE>(loc, sh, &tot), csum = 0.0;
#pragma unroll
    for (int k = 0; k < PEL; k++) if (i0 + k < N) { run += v[k]; cpre[p0 + k] = run; csum += run; }
    const double t = block_sum<WIDE>(csum, sh);
    if (tid == 0) part_c[blk] = t;
}

template <bool WIDE>
__device__ __forceinline__ void d_poisson_dsum(VrtFields F, const double* part_c, double* part_d, int G, const int blk) {
    using B = Blk<WIDE>;
    __shared__ double sh_[B::groups][40];
    double* sh = sh_[B::grp()];
    const int N = F.N, tid = B::tid(), i0 = blk * PTILE + tid * PEL, p0 = B::at(i0);
    const double* cpre = F.scratch + 2L * B::stride(N);
    double csum;
    partial_prefix<WIDE>(part_c, G, 0, sh, &csum);
    const double cmean = csum / (double)N;
    double loc = 0.0;
#pragma unroll
    for (int k = 0; k < PEL; k++) if (i0 + k < N) loc += (cmean - cpre[p0 + k]);
    const double t = block_sum<WIDE>(loc, sh);
    if (tid == 0) part_d[blk] = t;
}

template <bool WIDE>
__device__ __forceinline__ void d_poisson_scan2(VrtFields F, const double* part_c, const double* part_d, int G, const int blk, double* phi_s = nullptr) {
    using B = Blk<WIDE>;
    __shared__ double sh_[B::groups][40];
    double* sh = sh_[B::grp()];
    const int N = F.N, tid = B::tid(), i0 = blk * PTILE + tid * PEL, p0 = B::at(i0);
    const double* cpre = F.scratch + 2L * B::stride(N);
    double csum, tot;
    partial_prefix<WIDE>(part_c, G, 0, sh, &csum);
    const double cmean = csum / (double)N;
    const double base = partial_prefix<WIDE>(part_d, G, blk, sh, &tot);
    double d[PEL], loc = 0.0;
#pragma unroll
    for (int k = 0; k < PEL; k++) { d[k] = (i0 + k < N) ? (cmean - cpre[p0 + k]) : 0.0; loc += d[k]; }
    double run = base + block_excl_scan<WIDE>(loc, sh, &tot);
    const double sb = F.scratch[3L * B::stride(N)];
#pragma unroll
    for (int k = 0; k < PEL; k++) if (i0 + k < N) {
        const double phi = (i0 + k == 0) ? sb : run;
        F.PHI[i0 + k] = phi;
        if (WIDE) phi_s[p0 + k] = phi;           // the wide kernel tabulates E from this copy (padded layout, over the dead b vector)
        run += d[k];
    }
}


__global__ void __launch_bounds__(PTHR) k_poisson_rhs(VrtFields F, double* part) { vrt_pdl_sync(); d_poisson_rhs<false>(F, part, blockIdx.x); }
__global__ void __launch_bounds__(PTHR) k_poisson_conv(VrtFields F, const double* part_b, double* part_z, int G) { vrt_pdl_sync(); d_poisson_conv<false>(F, part_b, part_z, G, blockIdx.x); }
__global__ void __launch_bounds__(PTHR) k_poisson_scan1(VrtFields F, const double* part_z, double* part_c, int G) { vrt_pdl_sync(); d_poisson_scan1<false>(F, part_z, part_c, G, blockIdx.x); }
__global__ void __launch_bounds__(PTHR) k_poisson_dsum(VrtFields F, const double* part_c, double* part_d, int G) { vrt_pdl_sync(); d_poisson_dsum<false>(F, part_c, part_d, G, blockIdx.x); }
__global__ void __launch_bounds__(PTHR) k_poisson_scan2(VrtFields F, const double* part_c, const double* part_d, int G) { vrt_pdl_sync(); d_poisson_scan2<false>(F, part_c, part_d, G, blockIdx.x); }

// EMFieldSolver::GetEfield without the Ex0 term (EMSolver.cpp:137-154)
template <class Phi>
__device__ __forceinline__ double efield_of(const Phi& PHI, const int N, const double dx, int i) {
    int ip1 = i + 1, im1 = i - 1, ip2 = i + 2, im2 = i - 2;
    ip1 = ip1 > -1 ? ip1 : ip1 + N; im1 = im1 > -1 ? im1 : im1 + N;
    ip2 = ip2 > -1 ? ip2 : ip2 + N; im2 = im2 > -1 ? im2 : im2 + N;
    ip1 = ip1 < N ? ip1 : ip1 - N; im1 = im1 < N ? im1 : im1 - N;
    ip2 = ip2 < N ? ip2 : ip2 - N; im2 = im2 < N ? im2 : im2 - N;
    const double fieldCoef = 1.0 / (12 * dx);
    return -fieldCoef * (8 * (PHI(ip1) - PHI(im1)) - PHI(ip2) + PHI(im2));
}
__device__ __forceinline__ double efield_base(const VrtFields& F, int i) {
    return efield_of([&](int k) { return F.PHI[k]; }, F.N, F.dx, i);
}
// Ex0 += -(GetEfield(-1)+GetEfield(0))*0.5 (EMSolver.cpp:191, quirk Q4), then tabulate E on [-epad, N+epad)
__global__ void k_efield(VrtFields F, int update_ex0) { vrt_pdl_sync();
    __shared__ double ex0_new;
    if (threadIdx.x == 0) {
        double ex0 = *F.Ex0;
        if (update_ex0) ex0 += -((efield_base(F, -1) + ex0) + (efield_base(F, 0) + ex0)) * 0.5;
        ex0_new = ex0;
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x - F.epad;
    if (i < F.N + F.epad) F.E[i + F.epad] = efield_base(F, i) + ex0_new;
    if (blockIdx.x == 0 && threadIdx.x == 0) F.Ex0[1] = ex0_new;   // staged; committed by k_commit_ex0
}
__global__ void k_commit_ex0(VrtFields F) { vrt_pdl_sync(); F.Ex0[0] = F.Ex0[1]; }

// UpdatePotential + the E table in ONE launch for short x grids (G <= PSMALL tiles, i.e. N <= 4096: BASELINE configs 1, 2, 4, where a
// step is a chain of latency-bound launches and these seven are on its critical path): one CTA runs the same five passes over the
// tiles in turn — the same per-tile arithmetic and the same fixed-order sums of tile partials as the multi-CTA kernels, hence the
// same bits — then tabulates E and commits Ex0.
// The three N-vectors the passes hand to each other (b, z, prefix of z) and sum(b) live in shared memory here (3N + 1 doubles,
// <= 96 KB) instead of the global scratch: a single CTA would otherwise pay a global-memory round trip between every pair of passes.
__global__ void __launch_bounds__(PTHR) k_poisson_small(VrtFields F, double* part, int G) {
    vrt_pdl_sync();
    extern __shared__ __align__(16) double psm[];
    F.scratch = psm;
    part = psm + 3L * F.N + 8;            // the tile partials of the four passes too (PSMALL entries each)
    for (int blk = 0; blk < G; blk++) d_poisson_rhs<false>(F, part, blk);
    __syncthreads();
    for (int blk = 0; blk < G; blk++) { d_poisson_conv<false>(F, part, part + 8, G, blk); __syncthreads(); }
    for (int blk = 0; blk < G; blk++) { d_poisson_scan1<false>(F, part + 8, part + 16, G, blk); __syncthreads(); }
    for (int blk = 0; blk < G; blk++) { d_poisson_dsum<false>(F, part + 16, part + 24, G, blk); __syncthreads(); }
    for (int blk = 0; blk < G; blk++) { d_poisson_scan2<false>(F, part + 16, part + 24, G, blk); __syncthreads(); }
    __shared__ double ex0_new;
    if (threadIdx.x == 0) {
        const double ex0 = *F.Ex0;
        ex0_new = ex0 + -((efield_base(F, -1) + ex0) + (efield_base(F, 0) + ex0)) * 0.5;
    }
    __syncthreads();
    for (int i = (int)threadIdx.x - F.epad; i < F.N + F.epad; i += PTHR) F.E[i + F.epad] = efield_base(F, i) + ex0_new;
    if (threadIdx.x == 0) { F.Ex0[1] = ex0_new; F.Ex0[0] = ex0_new; }
}

// The tiles side by side inside ONE CTA: G groups of PTHR threads, one tile each (Blk<true>: named barriers inside a group, a
// CTA-wide barrier where the multi-CTA version has kernel boundaries), the three N-vectors and the tile partials in shared memory.
// Neither the global-memory round trips between the passes (cluster kernel) nor the walk over the tiles (single-block kernel).
__global__ void __launch_bounds__(PSMALL * PTHR) k_poisson_wide(VrtFields F, int G) {
    vrt_pdl_sync();
    extern __shared__ __align__(16) double psm[];
    F.scratch = psm;                                        // three padded N-vectors, sum(b), the tile partials, the groups' staged tiles
    double* part = psm + 3L * psw_len(F.N) + 8;
    double* tiles = part + 32;
    const double ex0 = *F.Ex0;                              // only this kernel writes Ex0: fetched now, used after the last pass
    const int blk = Blk<true>::grp();                       // blockDim.x = G * PTHR: every group has a tile
    d_poisson_rhs<true>(F, part, blk);
    __syncthreads();
    d_poisson_conv<true>(F, part, part + 8, G, blk, tiles);
    __syncthreads();
    d_poisson_scan1<true>(F, part + 8, part + 16, G, blk);
    __syncthreads();
    d_poisson_dsum<true>(F, part + 16, part + 24, G, blk);
    __syncthreads();
    d_poisson_scan2<true>(F, part + 16, part + 24, G, blk, psm);
    __syncthreads();
    __shared__ double ex0_new;
    const auto phi = [&](int k) { return psm[Blk<true>::at(k)]; };
    if (threadIdx.x == 0) ex0_new = ex0 + -((efield_of(phi, F.N, F.dx, -1) + ex0) + (efield_of(phi, F.N, F.dx, 0) + ex0)) * 0.5;
    __syncthreads();
    for (int i = (int)threadIdx.x - F.epad; i < F.N + F.epad; i += (int)blockDim.x) F.E[i + F.epad] = efield_of(phi, F.N, F.dx, i) + ex0_new;
    if (threadIdx.x == 0) { F.Ex0[1] = ex0_new; F.Ex0[0] = ex0_new; }
}

// The same solve with the tiles side by side: one thread-block cluster of up to PSMALL CTAs, one tile each, the hardware cluster barrier
// (release / acquire at cluster scope, so the global scratch one CTA wrote is visible to the others) where the multi-CTA version has
// kernel boundaries.  The passes run in parallel over the tiles instead of one after the other in a single CTA: the latency of five
// passes, not of five passes times the tile count, on the critical path of every RK stage of a short-grid step.  Same device
// functions, same tile index, same partial sums: same bits.
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(PTHR) k_poisson_cluster(VrtFields F, double* part, int G) {
    vrt_pdl_sync();
    const int blk = blockIdx.x, nc = gridDim.x;            // nc = cluster size (>= G, a power of two); CTAs beyond G only keep the barriers
    const bool work = blk < G;
    if (work) d_poisson_rhs<false>(F, part, blk);
    cluster_sync_all();
    if (work) d_poisson_conv<false>(F, part, part + PMAXT, G, blk);
    cluster_sync_all();
    if (work) d_poisson_scan1<false>(F, part + PMAXT, part + 2 * PMAXT, G, blk);
    cluster_sync_all();
    if (work) d_poisson_dsum<false>(F, part + 2 * PMAXT, part + 3 * PMAXT, G, blk);
    cluster_sync_all();
    if (work) d_poisson_scan2<false>(F, part + 2 * PMAXT, part + 3 * PMAXT, G, blk);
    cluster_sync_all();
    __shared__ double ex0_new;
    if (threadIdx.x == 0) {
        const double ex0 = *F.Ex0;
        ex0_new = ex0 + -((efield_base(F, -1) + ex0) + (efield_base(F, 0) + ex0)) * 0.5;
    }
    __syncthreads();
    for (int i = blk * PTHR + (int)threadIdx.x - F.epad; i < F.N + F.epad; i += nc * PTHR) F.E[i + F.epad] = efield_base(F, i) + ex0_new;
    cluster_sync_all();                                     // every CTA has read the old Ex0
    if (blk == 0 && threadIdx.x == 0) { F.Ex0[1] = ex0_new; F.Ex0[0] = ex0_new; }
}

// ---- EstimateCFLBound (EMSolver.cpp:631-664), un-offset indexing kept (quirk Q3) ---------------------
struct CflSpecies { int n; double m[8], q[8], dps[8]; double dpsMax; };
__global__ void k_cfl(VrtFields F, CflSpecies sp) {
    __shared__ double sh[32];
    double pc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F.N; i += gridDim.x * blockDim.x) {
        double Azv = F.Y[VRT_AZ][i], Ayv = F.Y[VRT_AY][i];
        double As = Ayv * Ayv + Azv * Azv, t = 0.0;
        for (int j = 0; j < sp.n; j++) t = fmax(t, fabs(sp.q[j]) / sp.m[j] / sqrt(1 + As / ((sp.m[j] * VRT_CS) * (sp.m[j] * VRT_CS))) * sp.dps[j]);
        pc = fmax(pc, t * fabs(Ayv * F.Y[VRT_BZ][i] - Azv * F.Y[VRT_BY][i]) + sp.dpsMax * fabs(F.E[i + F.epad]));
    }
    for (int o = 16; o > 0; o >>= 1) pc = fmax(pc, __shfl_down_sync(0xffffffffu, pc, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = pc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_down_sync(0xffffffffu, t, o));
        // pc >= 0: the IEEE bit pattern is monotone, so an integer max is exact
        if (threadIdx.x == 0) atomicMax((unsigned long long*)F.cfl, (unsigned long long)__double_as_longlong(t));
    }
}
__global__ void k_cfl_finish(VrtFields F) { double pc = *F.cfl; F.cfl[1] = 1.0 / fmax((VRT_CS / F.dx + pc), 1e-40); }

// ---- AssembleRhoAndJ bookkeeping (EMSolver.cpp:104-122, Level.cpp:19-62) -----------------------------
// EMFieldSolver::AssembleRhoAndJ (EMSolver.cpp:104-122) after the species' moment kernels, one thread per finest column: for every
// species in turn  charges[s][i] = 0 + (slab value | level sums from the finest level down, each level the sum over its patches in
// table order: Level::CollectRhoAndJ, Level.cpp:42-62, then Mesh::InterpolateRhoAndJToFinestMesh, Mesh.cpp:52-56),  J[i] carried
// across the species in the same order,  charge[i] = 0 + charges[0][i] + charges[1][i] + ...  — the additions of the reference's loops
// in their order, in one launch instead of three clears, one add per species and level set, and one total per species.
__global__ void k_assemble(VrtAssembleArgs A) { vrt_pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.N) return;
    double cu = 0.0, tot = 0.0;
    for (int s = 0; s < A.n_species; s++) {
        const VrtAssembleSpecies& S = A.sp[s];
        double ch = 0.0;
        if (S.mode == 0) {
            const int k = i - S.x0;
            if (k >= 0 && k < S.n) { ch += S.chargeR[k]; cu += S.currentR[k]; }
        } else {
            for (int d = 0; d < S.n_levels; d++) {
                if (!S.count[d]) continue;
                double lc = 0.0, lj = 0.0;
                for (int p = S.first[d]; p < S.first[d] + S.count[d]; p++) {
                    const VrtPatchDev& P = S.all[p];
                    const int k = i - P.x_pos * P.rtb;
                    if (k >= 0 && k < P.n_x * P.rtb) { lc += P.chargeR[k]; lj += P.currentR[k]; }
                }
                ch += lc; cu += lj;
            }
        }
        S.charges[i] = ch;
        tot += ch;
    }
    A.J[i] = cu;
    if (A.total) A.charge[i] = tot;
}
struct ChargePtrs { int n; const double* p[8]; };
__global__ void k_total_charge(ChargePtrs C, double* charge, int N) { vrt_pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double tot = 0.0;
    for (int s = 0; s < C.n; s++) tot += C.p[s][i];
    charge[i] = tot;
}
__global__ void k_neutralize(VrtFields F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F.N) { F.J[i] = 0.0; F.neutral[i] = -F.charge[i]; }
}

bool g_tab_loaded[64] = {};      // __constant__ memory is per device: one flag per device ordinal
inline int grid1(int n, int b = 256) { return (n + b - 1) / b; }

}  // namespace

int vrt_fields_init_tables(vrt_ctx* c) {
    bool& loaded = g_tab_loaded[c->device & 63];
    if (!loaded) { VRT_CUDA(c, cudaMemcpyToSymbol(c_tab, &kTableau, sizeof(VrtTableau))); loaded = true; }
    return 0;
}

// asq_out: where InterpolateToFaces writes a^2 at the x-faces (nullptr: F.a_squared).  Inside vrt_step the field stage runs
// concurrently with the Vlasov stage of the same RK stage, which still reads the previous a^2: it then writes the context's
// second buffer and the two are exchanged afterwards (enqueue_step, vrt_abi.cu).
int vrt_fields_rhs_update_faces(vrt_ctx* c, int step, const VrtStepParams* d_params, double* asq_out) {
    VrtFields F = c->F;
    if (asq_out) F.a_squared = asq_out;
    vrt_launch(k_field_rhs, dim3(grid1(F.M)), dim3(256), c->stream, F, step, d_params);
    vrt_launch(k_field_update, dim3(dim3(grid1(F.M), 6)), dim3(256), c->stream, F, step, d_params);
    vrt_launch(k_field_faces, dim3(grid1(F.N)), dim3(256), c->stream, F);
    c->launches += 3;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

int vrt_fields_poisson(vrt_ctx* c) {
    VrtFields& F = c->F;
    const int G = (F.N + PTILE - 1) / PTILE;
    if (G > PMAXT) { c->err = "vrt_poisson: x_size_finest too large for the tiled solver"; return VRT_ERR_ARG; }
    double* part = F.scratch + 3L * F.N + 8;      // 4 arrays of PMAXT tile partials behind the three N-vectors and sum(b)
    // short grids: VRT_POISSON_SMALL = 3 (default) one wide CTA with a thread group per tile, 2 one cluster launch, 1 one block walking the
    // tiles, 0 the multi-kernel passes (tests; profiles/ab_poisson_r2v.txt: 16 us against 25 us for the cluster at N = 2048)
    const int small_mode = getenv("VRT_POISSON_SMALL") ? atoi(getenv("VRT_POISSON_SMALL")) : 3;
    if (G <= PSMALL && small_mode == 3) {       // one CTA, the tiles side by side as thread groups
        const size_t smem = sizeof(double) * (3 * (size_t)psw_len(F.N) + 8 + 32 + (size_t)G * PTILE_SW);
        static bool attr_dev[64] = {};
        bool& attr = attr_dev[c->device & 63];
        if (!attr) { VRT_CUDA(c, cudaFuncSetAttribute(k_poisson_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (3 * psw_len(PSMALL * PTILE) + 8 + 32 + PSMALL * PTILE_SW)))); attr = true; }
        vrt_launch_smem(k_poisson_wide, dim3(1), dim3(G * PTHR), smem, c->stream, F, G);
        c->launches += 1;
        VRT_CUDA(c, cudaGetLastError());
        return 0;
    }
    if (G <= PSMALL && small_mode == 2) {
        unsigned nc = 1;
        while ((int)nc < G) nc <<= 1;
        vrt_launch_cluster(k_poisson_cluster, dim3(nc), dim3(PTHR), nc, c->stream, F, part, G);
        c->launches += 1;
        VRT_CUDA(c, cudaGetLastError());
        return 0;
    }
    if (G <= PSMALL && small_mode == 1) {
        const size_t smem = sizeof(double) * (3 * (size_t)F.N + 8 + 32);
        static bool attr_dev[64] = {};
        bool& attr = attr_dev[c->device & 63];
        if (!attr) { VRT_CUDA(c, cudaFuncSetAttribute(k_poisson_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (3 * PSMALL * PTILE + 8 + 32)))); attr = true; }
        vrt_launch_smem(k_poisson_small, dim3(1), dim3(PTHR), smem, c->stream, F, part, G);
        c->launches += 1;
        VRT_CUDA(c, cudaGetLastError());
        return 0;
    }
    vrt_launch(k_poisson_rhs, dim3(G), dim3(PTHR), c->stream, F, part);
    vrt_launch(k_poisson_conv, dim3(G), dim3(PTHR), c->stream, F, part, part + PMAXT, G);
    vrt_launch(k_poisson_scan1, dim3(G), dim3(PTHR), c->stream, F, part + PMAXT, part + 2 * PMAXT, G);
    vrt_launch(k_poisson_dsum, dim3(G), dim3(PTHR), c->stream, F, part + 2 * PMAXT, part + 3 * PMAXT, G);
    vrt_launch(k_poisson_scan2, dim3(G), dim3(PTHR), c->stream, F, part + 2 * PMAXT, part + 3 * PMAXT, G);
    c->launches += 4;
    vrt_launch(k_efield, dim3(grid1(F.N + 2 * F.epad)), dim3(256), c->stream, F, 1);
    vrt_launch(k_commit_ex0, dim3(1), dim3(1), c->stream, F);
    c->launches += 3;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// recompute the E table from the current PHI and Ex0 (after uploads)
int vrt_fields_refresh_efield(vrt_ctx* c) {
    VrtFields& F = c->F;
    vrt_launch(k_efield, dim3(grid1(F.N + 2 * F.epad)), dim3(256), c->stream, F, 0);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

int vrt_fields_cfl(vrt_ctx* c) {
    VrtFields& F = c->F;
    CflSpecies sp{}; sp.n = c->n_species; sp.dpsMax = 0.0;
    for (int i = 0; i < c->n_species; i++) {
        sp.m[i] = c->S[i].sp.m; sp.q[i] = c->S[i].sp.q; sp.dps[i] = 1 / c->S[i].sp.dp_finest;
        sp.dpsMax = fmax(sp.dpsMax, fabs(sp.q[i]) * sp.dps[i]);
    }
    VRT_CUDA(c, cudaMemsetAsync(F.cfl, 0, 2 * sizeof(double), c->stream));
    k_cfl<<<min(grid1(F.N), 148 * 4), 256, 0, c->stream>>>(F, sp);
    k_cfl_finish<<<1, 1, 0, c->stream>>>(F);
    c->launches += 2;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

int vrt_fields_assemble(vrt_ctx* c, unsigned mask, int total) {
    VrtAssembleArgs A{};
    A.N = c->F.N; A.total = total; A.J = c->F.J; A.charge = c->F.charge;
    for (int s = 0; s < c->n_species; s++) {
        if (!((mask >> s) & 1u)) continue;
        const VrtSpeciesState& S = c->S[s];
        VrtAssembleSpecies& T = A.sp[A.n_species++];
        T.charges = S.d_charges;
        if (S.path == VRT_PATH_FUSED) {
            T.mode = 0; T.chargeR = S.slab.chargeR; T.currentR = S.slab.currentR; T.x0 = S.slab.x_begin; T.n = S.slab.n_x;
        } else {
            T.mode = 1; T.all = S.d_patches; T.n_levels = (int)S.level_patches.size();
            if (T.n_levels > 16) { c->err = "vrt_moments: more than 16 levels"; return VRT_ERR_ARG; }
            for (int d = 0; d < T.n_levels; d++) { T.count[d] = (int)S.level_patches[d].size(); T.first[d] = T.count[d] ? S.level_patches[d][0] : 0; }
        }
    }
    vrt_launch(k_assemble, dim3(grid1(A.N)), dim3(256), c->stream, A);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}
int vrt_fields_total_charge(vrt_ctx* c) {
    ChargePtrs C{};
    C.n = c->n_species;
    for (int s = 0; s < c->n_species; s++) C.p[s] = c->S[s].d_charges;
    vrt_launch(k_total_charge, dim3(grid1(c->F.N)), dim3(256), c->stream, C, c->F.charge, c->F.N);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}
int vrt_fields_neutralize(vrt_ctx* c) {
    k_neutralize<<<grid1(c->F.N), 256, 0, c->stream>>>(c->F);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}
