// AMR machinery of the split path: everything Mesh::PushData (Mesh.cpp:91-106),
// Mesh::PushBoundaryC (Mesh.cpp:904-917) and the coarse-fine flux matching of Rectangle::FCTTimeStep sub-step 0
// (Rectangle.cpp:1313-1394) do between the sub-step kernels of vrt_split.cu.
//
//   host   vrt_conn_derive            Rectangle::CalculateConnectivitySame / FromFiner   Rectangle.cpp:671-864
//   K3     k_ghost_same               Rectangle::UpdateSameLevelBoundaries               Rectangle.cpp:562-614
//                                      (+ BoundaryCondition::GetValueFromSameLevel = 0.0, BoundaryCondition.cpp:6-8)
//   K5     k_restrict                 Rectangle::UpdateInterriorPoints                    Rectangle.cpp:314-337
//   K4     k_ghost_coarse             Rectangle::UpdateDifferentLevelBoundaries           Rectangle.cpp:343-560
//          k_corners                  Rectangle::UpdateCornerPoints                       Rectangle.cpp:1130-1214
//   K6     k_level_boundary_fluxes    CalculateFluxToCoarse* / RGKGetFlux*                Rectangle.hpp:132-178, Rectangle.cpp:1032-1253
//          k_boundary_c               Calculate{Same,Different}BoundaryC, UpdateSameBoundaryC   Rectangle.cpp:1625-1960
//
// One launch per level and pass; blockIdx.y selects the patch through the device patch table, threads run over the
// strips / ghost cells / faces of the patch perimeter (work is proportional to perimeter, except the restriction and the
// flagged-face scan which run over cells).  Compiled with -fmad=false: same operation order as the reference.
//
// Parallel execution is equivalent to the reference's serial patch loops because (i) same-level copies read neighbour
// interiors only (patch positions and sizes on refined levels are multiples of r, checked by vrt_conn_derive), (ii) every
// face of the limiter sync is written by at most two strips which compute the same value, (iii) levels are separate launches
// in the reference's order.
#include "vrt_internal.cuh"
#include "vrt_launch.cuh"
#include <cstring>
#include "vrt_device.cuh"
#include <algorithm>
#include <cmath>

// ===================================================================================================================
// host: connectivity
// ===================================================================================================================
namespace {

inline bool on_line(int x, int y1, int y2) { return ((x - y1) > -1) && ((y2 - x) > -1); }
inline long nsh(const vrt_patch_desc& q, int i, int j) { return (long)(q.n_p + 4) * (i + 2) + 2 + j; }

// coarse patch `ic` learns which of its cells/faces the finer patch `jf` covers; `jf` learns its coarser neighbours
void link_finer(vrt_conn& C, int ic, int jf) {
    const int r = C.r;
    const vrt_patch_desc &c = C.desc[ic], &f = C.desc[jf];
    VrtConnPatch &cc = C.P[ic], &cf = C.P[jf];
    const int fx0 = f.x_pos / r - c.x_pos, fp0 = f.p_pos / r - c.p_pos;      // finer patch in coarse-local cells
    const int fnx = f.n_x / r, fnp = f.n_p / r;
    const int fx1 = fx0 + fnx, fp1 = fp0 + fnp;
    const int lo_x = std::max(0, fx0), hi_x = std::min(c.n_x, fx1);
    const int lo_p = std::max(0, fp0), hi_p = std::min(c.n_p, fp1);
    for (int i = lo_x; i < hi_x; i++)
        for (int j = lo_p; j < hi_p; j++) { cc.flags[nsh(c, i, j)] |= VRT_NESTED; cc.finer[nsh(c, i, j)] = jf; }
    const int last = cf.ns_x - 1;
    auto corner_links = [&](int side) {
        if (on_line(fp0 - 1, 0, c.n_p - 1)) { cf.nb[side][0] = ic; cf.same[side][0] = 0; }
        if (on_line(fp1, 0, c.n_p - 1)) { cf.nb[side][last] = ic; cf.same[side][last] = 0; }
    };
    // upper x edge of the finer patch
    if (fx1 < c.n_x + 1 && fx1 > -1)
        for (int j = lo_p; j < hi_p; j++) { cc.finer_x[nsh(c, fx1, j)] = jf; cc.flags[nsh(c, fx1, j)] |= VRT_LBX; }
    if (fx1 < c.n_x && fx1 > -1) {
        for (int j = lo_p; j < hi_p; j++) { cf.nb[1][j - fp0 + 1] = ic; cf.same[1][j - fp0 + 1] = 0; }
        corner_links(1);
    }
    // lower x edge
    if (fx0 < c.n_x + 1 && fx0 > -1)
        for (int j = lo_p; j < hi_p; j++) { cc.finer_x[nsh(c, fx0, j)] = jf; cc.flags[nsh(c, fx0, j)] |= VRT_LBX; }
    if (fx0 < c.n_x + 1 && fx0 > 0) {
        for (int j = lo_p; j < hi_p; j++) { cf.nb[0][j - fp0 + 1] = ic; cf.same[0][j - fp0 + 1] = 0; }
        corner_links(0);
    }
    // upper / lower p edges
    if (fp1 < c.n_p + 1 && fp1 > -1)
        for (int i = lo_x; i < hi_x; i++) { cc.finer_p[nsh(c, i, fp1)] = jf; cc.flags[nsh(c, i, fp1)] |= VRT_LBP; }
    if (fp1 < c.n_p && fp1 > -1)
        for (int i = lo_x; i < hi_x; i++) { cf.nb[3][i - fx0] = ic; cf.same[3][i - fx0] = 0; }
    if (fp0 < c.n_p + 1 && fp0 > -1)
        for (int i = lo_x; i < hi_x; i++) { cc.finer_p[nsh(c, i, fp0)] = jf; cc.flags[nsh(c, i, fp0)] |= VRT_LBP; }
    if (fp0 < c.n_p + 1 && fp0 > 0)
        for (int i = lo_x; i < hi_x; i++) { cf.nb[2][i - fx0] = ic; cf.same[2][i - fx0] = 0; }
}

// patch `a` learns its same-level neighbour `b`
void link_same(vrt_conn& C, int a, int b) {
    const int r = C.r;
    const vrt_patch_desc &A = C.desc[a], &B = C.desc[b];
    VrtConnPatch& ca = C.P[a];
    const int rx = B.x_pos - A.x_pos, rp = B.p_pos - A.p_pos, last = ca.ns_x - 1;
    auto set = [&](int side, int e) { ca.nb[side][e] = b; ca.same[side][e] = 1; };
    if (rx == A.n_x) {                                  // b right of a
        if (on_line(-1, rp, rp + B.n_p - 1)) set(1, 0);
        if (on_line(A.n_p, rp, rp + B.n_p - 1)) set(1, last);
        for (int lo = std::max(B.p_pos, A.p_pos), hi = std::min(A.p_pos + A.n_p, B.p_pos + B.n_p); lo < hi; lo += r)
            set(1, (lo - A.p_pos) / r + 1);
    }
    if (rx == -B.n_x) {                                 // b left of a
        if (on_line(-1, rp, rp + B.n_p - 1)) set(0, 0);
        if (on_line(A.n_p, rp, rp + B.n_p - 1)) set(0, last);
        for (int lo = std::max(B.p_pos, A.p_pos), hi = std::min(A.p_pos + A.n_p, B.p_pos + B.n_p); lo < hi; lo += r)
            set(0, (lo - A.p_pos) / r + 1);
    }
    if (rp == A.n_p) {                                  // b above a
        if (on_line(-1, rx, rx + B.n_x - 1)) set(0, last);
        if (on_line(A.n_x, rx, rx + B.n_x - 1)) set(1, last);
        for (int lo = std::max(B.x_pos, A.x_pos), hi = std::min(A.x_pos + A.n_x, B.x_pos + B.n_x); lo < hi; lo += r)
            set(3, (lo - A.x_pos) / r);
    }
    if (rp == -B.n_p) {                                 // b below a
        if (on_line(-1, rx, rx + B.n_x - 1)) set(0, 0);
        if (on_line(A.n_x, rx, rx + B.n_x - 1)) set(1, 0);
        for (int lo = std::max(B.x_pos, A.x_pos), hi = std::min(A.x_pos + A.n_x, B.x_pos + B.n_x); lo < hi; lo += r)
            set(2, (lo - A.x_pos) / r);
    }
}

}  // namespace

// Connectivity passes of Mesh::promoteHierarchyToMesh (Mesh.cpp:840-861): finer->coarser links for every (coarse, fine)
// pair level by level, then same-level links; later links overwrite earlier ones, so patch order inside a level matters
// and is the caller's (= Level::rectangles order).
int vrt_conn_derive(vrt_conn& C, int n, const vrt_patch_desc* d, int r, int max_depth) {
    C.r = r; C.max_depth = max_depth;
    C.desc.assign(d, d + n);
    C.P.assign(n, VrtConnPatch());
    for (int p = 0; p < n; p++) {
        const vrt_patch_desc& q = d[p];
        if (q.depth < 0 || q.depth > max_depth || q.n_x < r || q.n_p < r || q.n_x % r || q.n_p % r) { C.err = "bad patch descriptor"; return VRT_ERR_ARG; }
        if (q.depth < max_depth && (q.x_pos % r || q.p_pos % r)) { C.err = "patches on refined levels must start on a multiple of the refinement ratio"; return VRT_ERR_ARG; }
        VrtConnPatch& cp = C.P[p];
        cp.ns_x = q.n_p / r + 2; cp.ns_p = q.n_x / r;
        const long npad = (long)(q.n_x + 4) * (q.n_p + 4);
        for (int s = 0; s < 4; s++) { cp.nb[s].assign(s < 2 ? cp.ns_x : cp.ns_p, -1); cp.same[s].assign(s < 2 ? cp.ns_x : cp.ns_p, 1); }
        cp.finer.assign(npad, -1); cp.finer_x.assign(npad, -1); cp.finer_p.assign(npad, -1);
        cp.flags.assign(npad, 0);
    }
    std::vector<std::vector<int>> lv(max_depth + 1);
    for (int p = 0; p < n; p++) lv[d[p].depth].push_back(p);
    for (int l = 1; l <= max_depth; l++)
        for (int a : lv[l]) for (int b : lv[l - 1]) link_finer(C, a, b);
    for (int l = 0; l <= max_depth; l++)
        for (int a : lv[l]) for (int b : lv[l]) if (a != b) link_same(C, a, b);
    return 0;
}

extern "C" {
int vrt_conn_create(vrt_conn** out, int n_patches, const vrt_patch_desc* patches, int refinement_ratio, int max_depth) {
    if (!out || n_patches < 1 || !patches || refinement_ratio < 2 || max_depth < 0) return VRT_ERR_ARG;
    vrt_conn* C = new vrt_conn();
    int rc = vrt_conn_derive(*C, n_patches, patches, refinement_ratio, max_depth);
    if (rc) { delete C; return rc; }
    *out = C;
    return 0;
}
int vrt_conn_strips(const vrt_conn* C, int patch, int side, int* nb, unsigned char* same) {
    if (!C || patch < 0 || patch >= (int)C->P.size() || side < 0 || side > 3) return VRT_ERR_ARG;
    const VrtConnPatch& p = C->P[patch];
    const int ns = (int)p.nb[side].size();
    for (int i = 0; i < ns; i++) { if (nb) nb[i] = p.nb[side][i]; if (same) same[i] = p.same[side][i]; }
    return ns;
}
int vrt_conn_flags(const vrt_conn* C, int patch, unsigned char* flags) {
    if (!C || patch < 0 || patch >= (int)C->P.size() || !flags) return VRT_ERR_ARG;
    std::copy(C->P[patch].flags.begin(), C->P[patch].flags.end(), flags);
    return 0;
}
void vrt_conn_destroy(vrt_conn* C) { delete C; }
}

// ===================================================================================================================
// device
// ===================================================================================================================
namespace {

// interpolationMatrix (Rectangle.cpp:76-89)
__constant__ double c_IMr[12] = {0.104166666666667, -0.708333333333334, 0.708333333333334, -0.104166666666667,
                                 0.117647058823529, 0.029411764705882,  0.029411764705882, 0.117647058823529,
                                 -0.083333333333333, 0.166666666666667, -0.166666666666667, 0.083333333333334};

__device__ __forceinline__ double* fstate(const VrtPatchDev& P, int val) { return val == 2 ? P.f2 : (val == 1 ? P.f1 : P.f0); }
// Rectangle::GetValueFromSameLevel (Rectangle.cpp:307-312); nb < 0: BoundaryCondition -> 0.0
__device__ __forceinline__ double same_level_value(const VrtPatchDev* all, int nb, int i, int j, int val) {
    if (nb < 0) return 0.0;
    const VrtPatchDev& Q = all[nb];
    return fstate(Q, val)[NS(Q, i - Q.x_pos, j - Q.p_pos)];
}
// Rectangle::GetInterpolantsREF (Rectangle.cpp:121-137) with the coefficients of Rectangle.cpp:94-101 evaluated in place.
// R = refinement ratio at compile time (0: run-time r).  With R = 2 — the only ratio the regrid path supports and the one every
// BASELINE config uses — the sub-cell loops unroll, the sub-cell coefficients fold to constants (the same IEEE operations, this unit is
// compiled with -fmad=false) and the small work arrays live in registers instead of local memory: the coarse -> fine ghost kernel,
// one thread per strip and therefore pure latency, was 56 % of the kernel time of a 3-level step before (profiles/launches_c4_r2e.txt).
constexpr int RMAX = 4;   // largest refinement ratio the interpolation buffers are sized for
template <int R>
__device__ __forceinline__ void interpolants_ref_t(int r, double f1, double f2, double f3, double f4, double f5, double* out) {
    const int rr = R ? R : r;
    f5 -= f3; f4 -= f3; f2 -= f3; f1 -= f3;
    double a1 = c_IMr[0] * f1 + c_IMr[1] * f2 + c_IMr[2] * f4 + c_IMr[3] * f5;
    double a2 = c_IMr[4] * f1 + c_IMr[5] * f2 + c_IMr[6] * f4 + c_IMr[7] * f5;
    double a3 = c_IMr[8] * f1 + c_IMr[9] * f2 + c_IMr[10] * f4 + c_IMr[11] * f5;
#pragma unroll
    for (int i = 0; i < (R ? R : RMAX); i++) {
        if (i >= rr) break;
        double tl = -0.5 + i / (double)rr, tr = -0.5 + (i + 1.0) / (double)rr;
        double c0 = (tl + tr) * 0.5;
        double c1 = (tl * tl + tl * tr + tr * tr) / 3.0 - (1.0 / 12);
        double c2 = (tl * tl * tl + tl * tl * tr + tl * tr * tr + tr * tr * tr) * 0.25;
        out[i] = c0 * a1 + c1 * a2 + c2 * a3 + f3;
    }
}
// the r x r sub-cells of the coarse cell of patch Q under fine cell (i, j) (fine-level global coordinates), with the mean-preserving
// correction: ip[k * r + l] = sub-cell (x k, p l) — Rectangle::GetWenoValueFromCoarseLevel (Rectangle.cpp:343-415)
template <int R>
__device__ __forceinline__ void coarse_block_t(const VrtPatchDev& Q, int r, int i, int j, int val, double* ip) {
    const int rr = R ? R : r;
    const double* f = fstate(Q, val);
    const int ic = i / rr - Q.x_pos, jc = j / rr - Q.p_pos;
    double temps[5][RMAX], part[RMAX], sum = 0.0;
#pragma unroll
    for (int k = -2; k < 3; k++)
        interpolants_ref_t<R>(r, f[NS(Q, ic - 2, jc + k)], f[NS(Q, ic - 1, jc + k)], f[NS(Q, ic, jc + k)], f[NS(Q, ic + 1, jc + k)], f[NS(Q, ic + 2, jc + k)], temps[k + 2]);
#pragma unroll
    for (int k = 0; k < (R ? R : RMAX); k++) {
        if (k >= rr) break;
        interpolants_ref_t<R>(r, temps[0][k], temps[1][k], temps[2][k], temps[3][k], temps[4][k], part);
#pragma unroll
        for (int l = 0; l < (R ? R : RMAX); l++) { if (l >= rr) break; ip[k * rr + l] = part[l]; sum += part[l]; }
    }
    const double correction = f[NS(Q, ic, jc)] - 1.0 / (double)(rr * rr) * sum;
#pragma unroll
    for (int k = 0; k < (R ? R * R : RMAX * RMAX); k++) { if (k >= rr * rr) break; ip[k] += correction; }
}
// out receives the two sub-cell layers nearest the fine patch (2r values) for side d = 0..3
template <int R>
__device__ __forceinline__ void coarse_values_t(const VrtPatchDev& Q, int r, int i, int j, int d, int val, double* out) {
    const int rr = R ? R : r;
    double ip[RMAX * RMAX];
    coarse_block_t<R>(Q, r, i, j, val, ip);
#pragma unroll
    for (int k = 0; k < (R ? R : RMAX); k++) {
        if (k >= rr) break;
        if (d == 0) { out[2 * k] = ip[rr * k + rr - 1]; out[2 * k + 1] = ip[rr * k + rr - 2]; }
        else if (d == 1) { out[2 * k] = ip[k]; out[2 * k + 1] = ip[rr + k]; }
        else if (d == 2) { out[2 * k] = ip[rr * k]; out[2 * k + 1] = ip[rr * k + 1]; }
        else { out[2 * k] = ip[rr * (rr - 1) + k]; out[2 * k + 1] = ip[rr * (rr - 2) + k]; }
    }
}
__device__ __forceinline__ void coarse_level_values(const VrtPatchDev& Q, int r, int i, int j, int d, int val, double* out) {
    if (r == 2) coarse_values_t<2>(Q, r, i, j, d, val, out); else coarse_values_t<0>(Q, r, i, j, d, val, out);
}
// the same interpolation, all r x r sub-cells (GetWenoValueFromCoarseLevel with d = -1, used by the regrid data transfer)
__device__ __forceinline__ void coarse_level_block(const VrtPatchDev& Q, int r, int i, int j, int val, double* ip) {
    if (r == 2) coarse_block_t<2>(Q, r, i, j, val, ip); else coarse_block_t<0>(Q, r, i, j, val, ip);
}

// ---- regrid data movers: Mesh::InterMeshDataTransfer (Mesh.cpp:116-130) ---------------------------
// `dst` = patches of one level of the NEW hierarchy (blockIdx.y), `src` = the n_src patches of one level of the old (or, for
// the last pass, new) hierarchy.  marks = Rectangle::is_interpolated of the new patches, one byte per padded cell, at
// mark_off[table index].
// Rectangle::GetDataFromSameLevelRectangle (Rectangle.cpp:920-941): every padded cell of the target that lies in the interior
// of a source patch of the same level takes its state-1 value (states 0 and 1), and is marked.  Source patches of a level are
// disjoint, so at most one matches.
__global__ void k_xfer_same(const VrtPatchDev* dst, int dst0, const VrtPatchDev* src, int n_src, unsigned char* marks, const long* mark_off) {
    const VrtPatchDev& P = dst[blockIdx.y];
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.npad) return;
    const int gi = P.x_pos + (int)(c / P.pitch) - 2, gj = P.p_pos + (int)(c % P.pitch) - 2;
    for (int q = 0; q < n_src; q++) {
        const VrtPatchDev& Q = src[q];
        if (gi >= Q.x_pos && gi < Q.x_pos + Q.n_x && gj >= Q.p_pos && gj < Q.p_pos + Q.n_p) {
            const double v = Q.f1[NS(Q, gi - Q.x_pos, gj - Q.p_pos)];
            P.f0[c] = v; P.f1[c] = v;
            marks[mark_off[dst0 + blockIdx.y] + c] = 1;
            return;
        }
    }
}
// Rectangle::GetDataFromCoarseLevelRectangle (Rectangle.cpp:892-918; only_unmarked = 0, marks the cells it writes) and
// Rectangle::GetDataFromCoarseNewLevelRectangle (Rectangle.cpp:1100-1128; only_unmarked = 1: skips coarse cells whose
// sub-cell (1,1) already holds old data, does not mark): one thread per coarse cell under the target plus one ring.
__global__ void k_xfer_coarse(const VrtPatchDev* dst, int dst0, const VrtPatchDev* src, int n_src, int r, int only_unmarked,
                              unsigned char* marks, const long* mark_off) {
    const VrtPatchDev& P = dst[blockIdx.y];
    const int tx = P.x_pos / r, tp = P.p_pos / r, tnx = P.n_x / r, tnp = P.n_p / r;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)(tnx + 2) * (tnp + 2)) return;
    const int ci = tx - 1 + (int)(t / (tnp + 2)), cj = tp - 1 + (int)(t % (tnp + 2));
    unsigned char* mk = marks + mark_off[dst0 + blockIdx.y];
    for (int q = 0; q < n_src; q++) {
        const VrtPatchDev& Q = src[q];
        if (ci >= Q.x_pos && ci < Q.x_pos + Q.n_x && cj >= Q.p_pos && cj < Q.p_pos + Q.n_p) {
            if (only_unmarked && mk[NS(P, r * (ci - tx) + 1, r * (cj - tp) + 1)]) return;
            double ip[RMAX * RMAX];
            coarse_level_block(Q, r, r * ci + 1, r * cj + 1, 1, ip);
            for (int k = 0; k < r; k++)
                for (int l = 0; l < r; l++) {
                    const long cell = NS(P, (ci - tx) * r + k, (cj - tp) * r + l);
                    P.f0[cell] = ip[l + r * k]; P.f1[cell] = ip[l + r * k];
                    if (!only_unmarked) mk[cell] = 1;
                }
            return;
        }
    }
}
// Rectangle::ErrorEstimate (Rectangle.hpp:128-130) on state 1, compared with the refinement criterion
struct ErrW { double w[5]; };
__global__ void k_error_flags(const VrtPatchDev* all, int patch, ErrW W, double criteria, unsigned char* out) {
    const VrtPatchDev& P = all[patch];
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)P.n_x * P.n_p) return;
    const int i = (int)(t / P.n_p), j = (int)(t % P.n_p);
    const double* f = P.f1;
    const double c = f[NS(P, i, j)], xp = f[NS(P, i + 1, j)], xm = f[NS(P, i - 1, j)], pp = f[NS(P, i, j + 1)], pm = f[NS(P, i, j - 1)];
    const double e = W.w[0] * fabs(xp - xm) + W.w[1] * fabs(pp - pm) + W.w[2] * fabs(xp - 2 * c + xm) + W.w[3] * fabs(pp - 2 * c + pm) + W.w[4] * fabs(c);
    out[t] = e > criteria ? 1 : 0;
}

// ---- K3: same-level ghost copy, side strips only (corners belong to k_corners / k_ghost_coarse) ---------------------
// thread t: [0, 2 n_p) cells of the xm / xp sides, [2 n_p, 2 n_p + 2 n_x) cells of the pm / pp sides; both ghost layers
__global__ void k_ghost_same(const VrtPatchDev* level, const VrtPatchDev* all, int val, int r) { vrt_pdl_sync();
    const VrtPatchDev& P = level[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = P.n_x, np = P.n_p;
    if (t >= 2 * np + 2 * nx) return;
    double* f = fstate(P, val);
    if (t < 2 * np) {
        const int side = t / np, j = t % np, e = j / r + 1;
        if (!P.same[side][e]) return;
        const int nb = P.nb[side][e];
        const int g1 = side == 0 ? -1 : nx, g2 = side == 0 ? -2 : nx + 1;
        f[NS(P, g1, j)] = same_level_value(all, nb, P.x_pos + g1, P.p_pos + j, val);
        f[NS(P, g2, j)] = same_level_value(all, nb, P.x_pos + g2, P.p_pos + j, val);
    } else {
        const int u = t - 2 * np, side = 2 + u / nx, i = u % nx, e = i / r;
        if (!P.same[side][e]) return;
        const int nb = P.nb[side][e];
        const int g1 = side == 2 ? -1 : np, g2 = side == 2 ? -2 : np + 1;
        f[NS(P, i, g1)] = same_level_value(all, nb, P.x_pos + i, P.p_pos + g1, val);
        f[NS(P, i, g2)] = same_level_value(all, nb, P.x_pos + i, P.p_pos + g2, val);
    }
}

// ---- K5: restriction of covered cells (Rectangle.cpp:314-337) -------------------------------------------------------
__global__ void k_restrict(const VrtPatchDev* level, const VrtPatchDev* all, int val, int r) { vrt_pdl_sync();
    const VrtPatchDev& P = level[blockIdx.y];
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.npad) return;
    if (!(P.flags[c] & VRT_NESTED)) return;
    const int i = (int)(c / P.pitch) - 2, j = (int)(c % P.pitch) - 2;
    const VrtPatchDev& Q = all[P.finer[c]];
    const double* fq = fstate(Q, val);
    const int i_f = (P.x_pos + i) * r - Q.x_pos, j_f = (P.p_pos + j) * r - Q.p_pos;
    double t = 0.0;
    for (int k = 0; k < r; k++) for (int l = 0; l < r; l++) t += fq[NS(Q, i_f + k, j_f + l)];
    t /= (double)(r * r);     // std::pow((double)refinementRatio, 2.0) is exact
    fstate(P, val)[c] = t;
}

// the four 2x2 ghost corners of a patch (Rectangle.cpp:1130-1214 = 484-560); c in [0, 8): side, which corner, p offset
__device__ void corner_cell(const VrtPatchDev& P, const VrtPatchDev* all, int val, int r, int c) {
    const int side = c >> 2, which = (c >> 1) & 1, jj = c & 1;
    const int nx = P.n_x;
    const int g1 = side == 0 ? -1 : nx, g2 = side == 0 ? -2 : nx + 1;
    const int i = which == 0 ? -1 : P.ns_x - 2;                 // strip index: -1 below the patch, n_p/r above it
    const int j = which == 0 ? r - 2 + jj : jj;                 // sub-cell inside the strip
    const int e = i + 1;
    const int nb = P.nb[side][e];
    double* f = fstate(P, val);
    if (!P.same[side][e]) {
        double temp[2 * RMAX];
        coarse_level_values(all[nb], r, P.x_pos + g1, P.p_pos + i * r, side == 0 ? 3 : 1, val, temp);
        f[NS(P, g1, i * r + j)] = temp[2 * j];
        f[NS(P, g2, i * r + j)] = temp[2 * j + 1];
    } else {
        f[NS(P, g1, i * r + j)] = same_level_value(all, nb, P.x_pos + g1, P.p_pos + i * r + j, val);
        f[NS(P, g2, i * r + j)] = same_level_value(all, nb, P.x_pos + g2, P.p_pos + i * r + j, val);
    }
}
__global__ void k_corners(const VrtPatchDev* level, const VrtPatchDev* all, int val, int r) { vrt_pdl_sync();
    if (threadIdx.x < 8) corner_cell(level[blockIdx.y], all, val, r, threadIdx.x);
}

// ---- K4: coarse -> fine ghost interpolation, one thread per strip with a coarser neighbour, then the corners ----------
__global__ void k_ghost_coarse(const VrtPatchDev* level, const VrtPatchDev* all, int val, int r) { vrt_pdl_sync();
    const VrtPatchDev& P = level[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int sx = P.ns_x - 2, sp = P.ns_p;
    if (t < 8) corner_cell(P, all, val, r, t);
    const int u = t - 8;
    if (u < 0 || u >= 2 * sx + 2 * sp) return;
    int side, i;
    if (u < 2 * sx) { side = u / sx; i = u % sx; } else { side = 2 + (u - 2 * sx) / sp; i = (u - 2 * sx) % sp; }
    const int e = side < 2 ? i + 1 : i;
    if (P.same[side][e]) return;
    const VrtPatchDev& Q = all[P.nb[side][e]];
    double* f = fstate(P, val);
    double temp[2 * RMAX];
    const int nx = P.n_x, np = P.n_p;
    if (side == 0) {
        coarse_level_values(Q, r, P.x_pos - 1, P.p_pos + i * r, 3, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, -1, i * r + j)] = temp[2 * j]; f[NS(P, -2, i * r + j)] = temp[2 * j + 1]; }
    } else if (side == 1) {
        coarse_level_values(Q, r, P.x_pos + nx, P.p_pos + i * r, 1, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, nx, i * r + j)] = temp[2 * j]; f[NS(P, nx + 1, i * r + j)] = temp[2 * j + 1]; }
    } else if (side == 2) {
        coarse_level_values(Q, r, P.x_pos + i * r, P.p_pos - 1, 0, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, i * r + j, -1)] = temp[2 * j]; f[NS(P, i * r + j, -2)] = temp[2 * j + 1]; }
    } else {
        coarse_level_values(Q, r, P.x_pos + i * r, P.p_pos + np, 2, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, i * r + j, np)] = temp[2 * j]; f[NS(P, i * r + j, np + 1)] = temp[2 * j + 1]; }
    }
}

// ---- K6a: coarse-fine flux matching ------------------------------------------------------------------------------------
// RGKGetFlux{X,P,XL,PL} (Rectangle.cpp:1032-1053, 1069-1098, 1216-1253) recomputed from the finer patch's f1;
// kind 0 X, 1 P, 2 XL, 3 PL.  Recursion depth <= number of levels.
__device__ double rgk_flux(const VrtPatchDev* all, int p, int i, int j, int kind, int r, const Sp& sp, const VrtFields& F);
__device__ double flux_to_coarse(const VrtPatchDev* all, int p, int i, int j, int kind, int r, const Sp& sp, const VrtFields& F) {
    const VrtPatchDev& P = all[p];
    i = i * r - P.x_pos; j = j * r - P.p_pos;
    double t = 0.0;
    for (int k = 0; k < r; k++) t += (kind == 0 || kind == 2) ? rgk_flux(all, p, i, j + k, kind, r, sp, F) : rgk_flux(all, p, i + k, j, kind, r, sp, F);
    t *= (1.0 / (double)(r * r));
    return t;
}
__device__ double rgk_flux(const VrtPatchDev* all, int p, int i, int j, int kind, int r, const Sp& sp, const VrtFields& F) {
    const VrtPatchDev& P = all[p];
    const long idx = NS(P, i, j);
    const bool isx = (kind == 0 || kind == 2);
    if (P.flags[idx] & (isx ? VRT_LBX : VRT_LBP))
        return flux_to_coarse(all, isx ? P.finer_x[idx] : P.finer_p[idx], P.x_pos + i, P.p_pos + j, kind, r, sp, F);
    const double w3 = 1.0 / 48.0, dx_inv = 1 / P.dx, dp_inv = 1 / P.dp, q = sp.q, cc = VRT_CS * VRT_CS * sp.m;
    const double* f = P.f1;
    const double Kp = __dmul_rn(dp_inv, cc), Kx = __dmul_rn(cc, dx_inv);
    if (isx) {
        const double as = q * q * a_sq(F, finest_index(P, i));
        const double g0 = gamma_(sp, momentum(P, sp, j), as), g1 = gamma_(sp, momentum(P, sp, j + 1), as);
        const double am = __dmul_rn(Kp, __dadd_rn(g1, -g0));
        if (kind == 2) return dx_inv * (am > 0.0 ? f[NS(P, i - 1, j)] : f[idx]) * am;
        const double g2 = gamma_(sp, momentum(P, sp, j + 2.0), as), gm = gamma_(sp, momentum(P, sp, j - 1.0), as);
        const double ap1 = __dmul_rn(Kp, __dadd_rn(g2, -g1)), am1 = __dmul_rn(Kp, __dadd_rn(g0, -gm));
        const double fm = weno(f[NS(P, i - 2, j)], f[NS(P, i - 1, j)], f[idx], f[NS(P, i + 1, j)], am > 0.0);
        const double fp1 = weno(f[NS(P, i - 2, j + 1)], f[NS(P, i - 1, j + 1)], f[idx + 1], f[NS(P, i + 1, j + 1)], ap1 > 0.0);
        const double fm1 = weno(f[NS(P, i - 2, j - 1)], f[NS(P, i - 1, j - 1)], f[idx - 1], f[NS(P, i + 1, j - 1)], am1 > 0.0);
        return dx_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
    }
    const double as_1 = q * q * a_sq(F, finest_index(P, i)), as_2 = q * q * a_sq(F, finest_index(P, i + 1));
    const double Em = q * patch_efield(P, F, i), mom = momentum(P, sp, j);
    const double G1 = gamma_(sp, mom, as_1), G2 = gamma_(sp, mom, as_2);
    const double am = __dadd_rn(Em, -__dmul_rn(Kx, __dadd_rn(G2, -G1)));
    if (kind == 3) return dp_inv * (am > 0.0 ? f[idx - 1] : f[idx]) * am;
    const double as_0 = q * q * a_sq(F, finest_index(P, i - 1)), as_3 = q * q * a_sq(F, finest_index(P, i + 2));
    const double Ep1 = q * patch_efield(P, F, i + 1), Em1 = q * patch_efield(P, F, i - 1);
    const double G0 = gamma_(sp, mom, as_0), G3 = gamma_(sp, mom, as_3);
    const double ap1 = __dadd_rn(Ep1, -__dmul_rn(Kx, __dadd_rn(G3, -G2)));
    const double am1 = __dadd_rn(Em1, -__dmul_rn(Kx, __dadd_rn(G1, -G0)));
    const long cp = idx + P.pitch, cm = idx - P.pitch;
    const double fm = weno(f[idx - 2], f[idx - 1], f[idx], f[idx + 1], am > 0.0);
    const double fp1 = weno(f[cp - 2], f[cp - 1], f[cp], f[cp + 1], ap1 > 0.0);
    const double fm1 = weno(f[cm - 2], f[cm - 1], f[cm], f[cm + 1], am1 > 0.0);
    return dp_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
}
// faces flagged is_interrior_level_boundary_{x,p}: flux := mean of the finer patch's face fluxes (Rectangle.cpp:1313-1394)
__global__ void k_level_boundary_fluxes(const VrtPatchDev* level, const VrtPatchDev* all, int step, int r, Sp sp, VrtFields F) { vrt_pdl_sync();
    const VrtPatchDev& P = level[blockIdx.y];
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.npad) return;
    const unsigned char fl = P.flags[c];
    if (!(fl & (VRT_LBX | VRT_LBP))) return;
    const int i = (int)(c / P.pitch) - 2, j = (int)(c % P.pitch) - 2;
    const int nx = P.n_x, np = P.n_p;
    if ((fl & VRT_LBX) && i >= 0 && i <= nx && j >= -1 && j <= np) {
        P.FxH[step * P.npad + c] = flux_to_coarse(all, P.finer_x[c], P.x_pos + i, P.p_pos + j, 0, r, sp, F);
        if (step == 0) P.FxL[c] = flux_to_coarse(all, P.finer_x[c], P.x_pos + i, P.p_pos + j, 2, r, sp, F);
    }
    if ((fl & VRT_LBP) && i >= -1 && i <= nx && j >= 0 && j <= np) {
        P.FpH[step * P.npad + c] = flux_to_coarse(all, P.finer_p[c], P.x_pos + i, P.p_pos + j, 1, r, sp, F);
        if (step == 0) P.FpL[c] = flux_to_coarse(all, P.finer_p[c], P.x_pos + i, P.p_pos + j, 3, r, sp, F);
    }
}

// The same over the list of flagged faces (refinement ratio 2): one thread per (face, flux kind, fine sub-face), so that a warp is
// 32 lanes of equal work instead of the one or two flagged cells a warp of the scan above holds; the two fine-face fluxes of a coarse
// face meet by shuffle and are added in the reference's order  t = 0 + F(k = 0) + F(k = 1), t *= 1/r^2  (Rectangle.hpp:132-178).
__global__ void k_lb_fluxes_r2(const VrtLbFace* faces, int n, const VrtPatchDev* all, int step, Sp sp, VrtFields F) { vrt_pdl_sync();
    constexpr int r = 2;
    const int kinds = step == 0 ? 2 : 1;              // high-order always, low-order at stage 0 only (quirk Q1)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = t / (kinds * r), rem = t % (kinds * r), low = rem / r, k = rem % r;
    const bool valid = e < n;
    double val = 0.0;
    int dir = 0; long c = 0; int patch = 0;
    if (valid) {
        const VrtLbFace f = faces[e];
        patch = f.patch; dir = f.dir; c = f.cell;
        const VrtPatchDev& P = all[patch];
        const int i = (int)(c / P.pitch) - 2, j = (int)(c % P.pitch) - 2;
        const int finer = dir == 0 ? P.finer_x[c] : P.finer_p[c];
        const VrtPatchDev& Q = all[finer];
        const int ii = (P.x_pos + i) * r - Q.x_pos, jj = (P.p_pos + j) * r - Q.p_pos;
        const int kind = dir + 2 * low;               // 0 X, 1 P, 2 XL, 3 PL
        val = dir == 0 ? rgk_flux(all, finer, ii, jj + k, kind, r, sp, F) : rgk_flux(all, finer, ii + k, jj, kind, r, sp, F);
    }
    const double other = __shfl_down_sync(0xffffffffu, val, 1);      // lane k = 0 receives the k = 1 flux of the same face and kind
    if (valid && k == 0) {
        double tsum = 0.0;
        tsum += val; tsum += other;
        tsum *= (1.0 / (double)(r * r));
        const VrtPatchDev& P = all[patch];
        if (dir == 0) { if (low) P.FxL[c] = tsum; else P.FxH[step * P.npad + c] = tsum; }
        else { if (low) P.FpL[c] = tsum; else P.FpH[step * P.npad + c] = tsum; }
    }
}

// ---- K6b: limiter sync, one thread per side strip ------------------------------------------------------------------------
// pass 4: CalculateSameBoundaryC -> SetCFromSameLevel (Rectangle.cpp:1835-1877, 1795-1833)
// pass 5: CalculateDifferentBoundaryC -> SetCFromDifferentLevel (1706-1763, 1625-1704)
// pass 6: UpdateSameBoundaryC -> UpdateCFromSameLevel (1919-1960, 1879-1917)
__global__ void k_boundary_c(const VrtPatchDev* level, const VrtPatchDev* all, int pass, int r) { vrt_pdl_sync();
    const VrtPatchDev& Cl = level[blockIdx.y];       // the caller
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int sx = Cl.ns_x - 2, sp = Cl.ns_p;
    if (u >= 2 * sx + 2 * sp) return;
    int t, s;
    if (u < 2 * sx) { t = u / sx; s = u % sx; } else { t = 2 + (u - 2 * sx) / sp; s = (u - 2 * sx) % sp; }
    const int e = t < 2 ? s + 1 : s;
    const bool same = Cl.same[t][e];
    const int nb = Cl.nb[t][e];
    if ((pass == 5) == same) return;
    if (nb < 0) return;                            // BoundaryCondition: zero-trip loops (quirk Q8)
    const VrtPatchDev& N = all[nb];
    // global coordinates of the first face of the strip
    const int gi = t == 1 ? Cl.x_pos + Cl.n_x : (t < 2 ? Cl.x_pos : Cl.x_pos + r * s);
    const int gj = t == 3 ? Cl.p_pos + Cl.n_p : (t < 2 ? Cl.p_pos + r * s : Cl.p_pos);
    const int icl = gi - Cl.x_pos, jcl = gj - Cl.p_pos;
    const bool xdir = t < 2;
    // cell of the caller adjacent to the face on the caller's side (t = 0, 2: the face's own cell; t = 1, 3: one below)
    const int ax = (t == 1) ? -1 : 0, ap = (t == 3) ? -1 : 0;
    if (pass == 5) {
        const int ico = gi / r - N.x_pos, jco = gj / r - N.p_pos;
        const long nc = NS(N, ico, jco);
        double cv = xdir ? N.Cx[nc] : N.Cp[nc];
        for (int k = 0; k < r; k++) {
            const long fc = xdir ? NS(Cl, icl, jcl + k) : NS(Cl, icl + k, jcl);
            const long ac = xdir ? NS(Cl, icl + ax, jcl + k) : NS(Cl, icl + k, jcl + ap);
            const double fds = xdir ? Cl.FxDS[fc] : Cl.FpDS[fc];
            // t = 0, 2: the fine cell receives a positive flux (Rp); t = 1, 3: it loses it (Rm)
            const bool recv = (t == 0 || t == 2);
            cv = fmin(cv, (fds > 0.0) == recv ? Cl.Rp[ac] : Cl.Rm[ac]);
        }
        {   // the coarse cell on the far side of the face
            const long oc = (t == 0) ? NS(N, ico - 1, jco) : ((t == 2) ? NS(N, ico, jco - 1) : nc);
            const double fds = xdir ? N.FxDS[nc] : N.FpDS[nc];
            const bool recv = (t == 1 || t == 3);   // t = 1, 3: the coarse cell is the face's own cell
            cv = fmin(cv, (fds > 0.0) == recv ? N.Rp[oc] : N.Rm[oc]);
        }
        for (int k = 0; k < r; k++) {
            const long fc = xdir ? NS(Cl, icl, jcl + k) : NS(Cl, icl + k, jcl);
            if (xdir) Cl.Cx[fc] = cv; else Cl.Cp[fc] = cv;
        }
        if (xdir) N.Cx[nc] = cv; else N.Cp[nc] = cv;
        return;
    }
    const int in = gi - N.x_pos, jn = gj - N.p_pos;
    for (int k = 0; k < r; k++) {
        const long fn = xdir ? NS(N, in, jn + k) : NS(N, in + k, jn);        // the face in the neighbour
        const long fc = xdir ? NS(Cl, icl, jcl + k) : NS(Cl, icl + k, jcl);  // the face in the caller
        double cv;
        if (pass == 4) {
            // the two cells sharing the face: `lo` below it, `hi` above it; the caller holds `hi` for t = 0, 2
            const long ac = xdir ? NS(Cl, icl + ax, jcl + k) : NS(Cl, icl + k, jcl + ap);
            const long an = (t == 0) ? NS(N, in - 1, jn + k) : ((t == 2) ? NS(N, in + k, jn - 1) : fn);
            const double fds = xdir ? N.FxDS[fn] : N.FpDS[fn];
            const bool caller_hi = (t == 0 || t == 2);
            const double Rp_hi = caller_hi ? Cl.Rp[ac] : N.Rp[an], Rm_hi = caller_hi ? Cl.Rm[ac] : N.Rm[an];
            const double Rp_lo = caller_hi ? N.Rp[an] : Cl.Rp[ac], Rm_lo = caller_hi ? N.Rm[an] : Cl.Rm[ac];
            cv = fds > 0.0 ? fmin(Rp_hi, Rm_lo) : fmin(Rp_lo, Rm_hi);
        } else {
            cv = xdir ? fmin(N.Cx[fn], Cl.Cx[fc]) : fmin(N.Cp[fn], Cl.Cp[fc]);
        }
        if (xdir) { N.Cx[fn] = cv; Cl.Cx[fc] = cv; } else { N.Cp[fn] = cv; Cl.Cp[fc] = cv; }
    }
}

inline Sp make_sp(const VrtSpecies& s) { return Sp{s.m, s.q, s.pmin, 1 / s.m}; }

}  // namespace

// ===================================================================================================================
// host launchers
// ===================================================================================================================
int vrt_amr_upload_connectivity(vrt_ctx* c, int s, const vrt_conn& C) {
    VrtSpeciesState& S = c->S[s];
    const int n = (int)C.P.size();
    // pool layout per patch (table order): nb[4] ints, finer/finer_x/finer_p ints, then bytes same[4], flags
    size_t bytes = 0;
    std::vector<size_t> off(n);
    for (int ti = 0; ti < n; ti++) {
        const VrtConnPatch& p = C.P[S.table_order[ti]];
        off[ti] = bytes;
        size_t ints = 2 * (size_t)p.ns_x + 2 * (size_t)p.ns_p + 3 * p.flags.size();
        size_t chars = 2 * (size_t)p.ns_x + 2 * (size_t)p.ns_p + p.flags.size();
        bytes += ints * sizeof(int) + ((chars + 15) / 16) * 16;
    }
    // the flagged coarse faces (index ranges of Rectangle.cpp:1313-1394: x-faces i in [0, n_x], j in [-1, n_p]; p-faces i in [-1, n_x],
    // j in [0, n_p]), in table order, hence grouped by depth
    std::vector<VrtLbFace> lb;
    S.lb_first.assign(S.level_patches.size(), 0); S.lb_count.assign(S.level_patches.size(), 0);
    for (int ti = 0; ti < n; ti++) {
        const VrtConnPatch& p = C.P[S.table_order[ti]];
        const VrtPatchDev& T = S.table[ti];
        const size_t before = lb.size();
        for (long cidx = 0; cidx < T.npad; cidx++) {
            const unsigned char fl = p.flags[cidx];
            if (!(fl & (VRT_LBX | VRT_LBP))) continue;
            const int i = (int)(cidx / T.pitch) - 2, j = (int)(cidx % T.pitch) - 2;
            if ((fl & VRT_LBX) && i >= 0 && i <= T.n_x && j >= -1 && j <= T.n_p) lb.push_back(VrtLbFace{ti, 0, cidx});
            if ((fl & VRT_LBP) && i >= -1 && i <= T.n_x && j >= 0 && j <= T.n_p) lb.push_back(VrtLbFace{ti, 1, cidx});
        }
        if (lb.size() > before) {
            if (S.lb_count[T.depth] == 0) S.lb_first[T.depth] = (int)before;
            S.lb_count[T.depth] += (int)(lb.size() - before);
        }
    }
    bytes = (bytes + 15) & ~(size_t)15;
    const size_t lb_off = bytes;
    bytes += lb.size() * sizeof(VrtLbFace);
    std::vector<unsigned char> host(bytes, 0);
    if (!lb.empty()) std::memcpy(host.data() + lb_off, lb.data(), lb.size() * sizeof(VrtLbFace));
    if (S.conn_pool) { cudaFree(S.conn_pool); S.conn_pool = nullptr; }
    VRT_CUDA(c, cudaMalloc(&S.conn_pool, std::max<size_t>(bytes, 16)));
    unsigned char* dbase = (unsigned char*)S.conn_pool;
    S.d_lb = lb.empty() ? nullptr : (VrtLbFace*)(dbase + lb_off);
    S.has_amr = false;
    for (int ti = 0; ti < n; ti++) {
        const VrtConnPatch& p = C.P[S.table_order[ti]];
        VrtPatchDev& T = S.table[ti];
        T.ns_x = p.ns_x; T.ns_p = p.ns_p;
        int* hi = (int*)(host.data() + off[ti]);
        int* di = (int*)(dbase + off[ti]);
        size_t k = 0;
        auto to_table = [&](int caller_idx) { return caller_idx < 0 ? -1 : S.table_index[caller_idx]; };
        for (int sd = 0; sd < 4; sd++) {
            T.nb[sd] = di + k;
            for (int v : p.nb[sd]) { hi[k++] = to_table(v); if (v >= 0) S.has_amr = true; }
        }
        const std::vector<int>* planes[3] = {&p.finer, &p.finer_x, &p.finer_p};
        int** dst[3] = {&T.finer, &T.finer_x, &T.finer_p};
        for (int q = 0; q < 3; q++) {
            *dst[q] = di + k;
            for (int v : *planes[q]) hi[k++] = to_table(v);
        }
        unsigned char* hc = (unsigned char*)(hi + k);
        unsigned char* dc = (unsigned char*)(di + k);
        size_t m = 0;
        for (int sd = 0; sd < 4; sd++) {
            T.same[sd] = dc + m;
            for (unsigned char v : p.same[sd]) hc[m++] = v;
        }
        T.flags = dc + m;
        for (unsigned char v : p.flags) { hc[m++] = v; if (v) S.has_amr = true; }
    }
    VRT_CUDA(c, cudaMemcpyAsync(S.conn_pool, host.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t p = 0; p < S.patches.size(); p++) S.patches[p] = S.table[S.table_index[p]];
    return 0;
}

static unsigned blocks(long n, int b = 128) { return (unsigned)((n + b - 1) / b); }

// Level::PushData(updateType, val) (Level.cpp:88-126): 0 restriction, 1 same-level strips, 2 coarser-level strips + corners,
// 3 corners, 4 / 5 / 6 the three limiter-sync passes.  l = -1: every level in one launch — valid for the passes that only touch
// patches of one level (1, 4, 6): inside such a pass the levels do not interact, and on a small hierarchy every launch saved is
// several microseconds of the step.
int vrt_amr_level_pass(vrt_ctx* c, int s, int l, int type, int val) {
    VrtSpeciesState& S = c->S[s];
    const int nl = (int)S.level_patches.size(), r = c->refinement_ratio;
    if (l < -1 || l >= nl || type < 0 || type > 6) { c->err = "level pass: bad level or update type"; return VRT_ERR_ARG; }
    if (l == -1 && !(type == 1 || type == 4 || type == 6)) { c->err = "level pass: this pass couples levels and runs level by level"; return VRT_ERR_ARG; }
    const VrtPatchDev* all = S.d_patches;
    int first = 0, count = (int)S.table.size();
    if (l >= 0) {
        const std::vector<int>& lp = S.level_patches[l];
        if (lp.empty()) return 0;
        first = lp[0]; count = (int)lp.size();
    }
    if (count == 0) return 0;
    const VrtPatchDev* level = all + first;
    const unsigned np = (unsigned)count;
    long perim = 0, strips = 0, npad = 0;
    for (int p = first; p < first + count; p++) {
        const VrtPatchDev& T = S.table[p];
        perim = std::max<long>(perim, 2L * T.n_x + 2L * T.n_p);
        strips = std::max<long>(strips, 2L * (T.ns_x - 2) + 2L * T.ns_p);
        npad = std::max(npad, T.npad);
    }
    switch (type) {
        case 0:
            if (l == 0 || S.level_patches[l - 1].empty()) return 0;      // nothing is nested without a finer level
            vrt_launch(k_restrict, dim3(dim3(blocks(npad, 256), np)), dim3(256), c->stream, level, all, val, r); break;
        case 1: vrt_launch(k_ghost_same, dim3(dim3(blocks(perim), np)), dim3(128), c->stream, level, all, val, r); break;
        case 2: vrt_launch(k_ghost_coarse, dim3(dim3(blocks(strips + 8), np)), dim3(128), c->stream, level, all, val, r); break;
        case 3: vrt_launch(k_corners, dim3(dim3(1, np)), dim3(32), c->stream, level, all, val, r); break;
        default:
            if (!S.has_amr) return 0;     // every strip faces the BoundaryCondition object: zero-trip loops (quirk Q8)
            vrt_launch(k_boundary_c, dim3(dim3(blocks(strips), np)), dim3(128), c->stream, level, all, type, r); break;
    }
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// Mesh::PushData(val) (Mesh.cpp:91-106): finest level's same-level copy, then per coarser level restriction + same-level copy,
// the coarsest level's corners, then the coarse -> fine ghosts from coarse to fine.  Reordered without changing a value: the
// restrictions read fine-patch interiors only (not the ghost layers the same-level copies write) and each level's same-level copy
// reads interiors of its own level only — so all restrictions go first, level by level (a level's restriction averages cells the
// finer level's restriction wrote), and the same-level copies of all levels follow in ONE launch.
int vrt_amr_push_data(vrt_ctx* c, int s, int val) {
    const int nl = (int)c->S[s].level_patches.size();
    int rc;
    for (int l = 1; l < nl; l++) if ((rc = vrt_amr_level_pass(c, s, l, 0, val))) return rc;
    if ((rc = vrt_amr_level_pass(c, s, -1, 1, val))) return rc;
    if ((rc = vrt_amr_level_pass(c, s, nl - 1, 3, val))) return rc;
    for (int l = nl - 1; l > 0; l--) if ((rc = vrt_amr_level_pass(c, s, l - 1, 2, val))) return rc;
    return 0;
}

// Mesh::PushBoundaryC (Mesh.cpp:904-917): passes 4, 5, 6 over all levels, finest first.  Passes 4 and 6 pair patches of one
// level only: one launch each for all levels; pass 5 writes the coarse face next to a fine strip and stays level by level.
int vrt_amr_push_boundary_c(vrt_ctx* c, int s) {
    const int nl = (int)c->S[s].level_patches.size();
    if (int rc = vrt_amr_level_pass(c, s, -1, 4, 1)) return rc;
    // (the coarsest level has no coarser neighbour: its pass 5 would find no strip to work on)
    for (int l = 0; l + 1 < nl; l++) if (int rc = vrt_amr_level_pass(c, s, l, 5, 1)) return rc;
    return vrt_amr_level_pass(c, s, -1, 6, 1);
}

static int launch_lb_list(vrt_ctx* c, VrtSpeciesState& S, int first, int count, int step) {
    const int kinds = step == 0 ? 2 : 1;
    const long threads = (long)count * kinds * 2;
    vrt_launch(k_lb_fluxes_r2, dim3(blocks(threads, 128)), dim3(128), c->stream, (const VrtLbFace*)(S.d_lb + first), count, (const VrtPatchDev*)S.d_patches, step,
               make_sp(S.sp), c->F);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

int vrt_amr_level_boundary_fluxes(vrt_ctx* c, int s, int depth, int step) {
    VrtSpeciesState& S = c->S[s];
    if (depth == 0 || S.level_patches[depth].empty() || S.level_patches[depth - 1].empty()) return 0;
    if (c->refinement_ratio == 2 && S.d_lb) return S.lb_count[depth] ? launch_lb_list(c, S, S.lb_first[depth], S.lb_count[depth], step) : 0;
    long m = 0;
    for (int p : S.level_patches[depth]) m = std::max(m, S.table[p].npad);
    vrt_launch(k_level_boundary_fluxes, dim3(dim3(blocks(m, 128), (unsigned)S.level_patches[depth].size())), dim3(128), c->stream, 
        S.d_patches + S.level_patches[depth][0], S.d_patches, step, c->refinement_ratio, make_sp(S.sp), c->F);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// the same for every coarse level in one launch: the table is grouped by depth (finest first), so the patches of depth >= 1
// are its tail; patches without a finer level carry no flagged faces and fall through
int vrt_amr_level_boundary_fluxes_all(vrt_ctx* c, int s, int step) {
    VrtSpeciesState& S = c->S[s];
    int first = -1;
    long m = 0;
    for (size_t d = 1; d < S.level_patches.size(); d++)
        for (int p : S.level_patches[d]) { if (first < 0) first = p; m = std::max(m, S.table[p].npad); }
    if (first < 0 || !S.has_amr) return 0;
    if (c->refinement_ratio == 2 && S.d_lb) {
        int lo = -1, cnt = 0;
        for (size_t d = 1; d < S.lb_count.size(); d++) if (S.lb_count[d]) { if (lo < 0) lo = S.lb_first[d]; cnt += S.lb_count[d]; }
        return cnt ? launch_lb_list(c, S, lo, cnt, step) : 0;
    }
    vrt_launch(k_level_boundary_fluxes, dim3(dim3(blocks(m, 128), (unsigned)(S.table.size() - first))), dim3(128), c->stream, 
        S.d_patches + first, S.d_patches, step, c->refinement_ratio, make_sp(S.sp), c->F);
    c->launches += 1;
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

// Mesh::InterMeshDataTransfer (Mesh.cpp:116-130) between the old and the new device patch tables of one species
int vrt_amr_transfer(vrt_ctx* c, const VrtSpeciesState& O, VrtSpeciesState& N) {
    const int r = c->refinement_ratio, nl = (int)N.level_patches.size();
    if (r > RMAX) { c->err = "vrt_regrid: refinement ratio too large"; return VRT_ERR_ARG; }
    // Rectangle::is_interpolated of the new patches
    std::vector<long> off(N.table.size() + 1, 0);
    for (size_t k = 0; k < N.table.size(); k++) off[k + 1] = off[k] + N.table[k].npad;
    unsigned char* marks = nullptr; long* d_off = nullptr;
    VRT_CUDA(c, cudaMalloc(&marks, std::max<long>(off.back(), 1)));
    VRT_CUDA(c, cudaMalloc(&d_off, sizeof(long) * off.size()));
    VRT_CUDA(c, cudaMemsetAsync(marks, 0, std::max<long>(off.back(), 1), c->stream));
    VRT_CUDA(c, cudaMemcpyAsync(d_off, off.data(), sizeof(long) * off.size(), cudaMemcpyHostToDevice, c->stream));
    auto level_of = [](const VrtSpeciesState& S, int l, const VrtPatchDev** first, int* n) {
        *n = (l >= 0 && l < (int)S.level_patches.size()) ? (int)S.level_patches[l].size() : 0;
        *first = *n ? S.d_patches + S.level_patches[l][0] : nullptr;
    };
    auto same = [&](int l) {          // new level l <- old level l
        const VrtPatchDev *dst, *src; int nd, ns;
        level_of(N, l, &dst, &nd); level_of(O, l, &src, &ns);
        if (!nd || !ns) return;
        long npad = 0;
        for (int p : N.level_patches[l]) npad = std::max(npad, N.table[p].npad);
        k_xfer_same<<<dim3(blocks(npad, 256), nd), 256, 0, c->stream>>>(dst, N.level_patches[l][0], src, ns, marks, d_off);
        c->launches += 1;
    };
    auto coarse = [&](int l, const VrtSpeciesState& Src, int only_unmarked) {      // new level l <- level l + 1 of Src
        const VrtPatchDev *dst, *src; int nd, ns;
        level_of(N, l, &dst, &nd); level_of(Src, l + 1, &src, &ns);
        if (!nd || !ns) return;
        long cells = 0;
        for (int p : N.level_patches[l]) cells = std::max(cells, (long)(N.table[p].n_x / r + 2) * (N.table[p].n_p / r + 2));
        k_xfer_coarse<<<dim3(blocks(cells, 128), nd), 128, 0, c->stream>>>(dst, N.level_patches[l][0], src, ns, r, only_unmarked, marks, d_off);
        c->launches += 1;
    };
    for (int l = 0; l + 1 < nl; l++) { coarse(l, O, 0); same(l); }
    if (nl > 1) same(nl - 1);
    for (int l = nl - 1; l > 0; l--) coarse(l - 1, N, 1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(marks); cudaFree(d_off);
    if (e != cudaSuccess) { c->err = std::string("vrt_regrid transfer: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
    return 0;
}

int vrt_amr_error_flags(vrt_ctx* c, int s, int patch, const double weights[5], double criteria, unsigned char* flags_host) {
    VrtSpeciesState& S = c->S[s];
    const VrtPatchDev& P = S.patches[patch];
    const long n = (long)P.n_x * P.n_p;
    unsigned char* d = nullptr;
    VRT_CUDA(c, cudaMalloc(&d, n));
    ErrW W; for (int k = 0; k < 5; k++) W.w[k] = weights[k];
    k_error_flags<<<blocks(n, 256), 256, 0, c->stream>>>(S.d_patches, S.table_index[patch], W, criteria, d);
    c->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags_host, d, n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) { c->err = std::string("vrt_error_flags: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
    return 0;
}
