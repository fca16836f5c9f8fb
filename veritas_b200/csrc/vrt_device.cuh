// Device helpers shared by the split and fused Vlasov kernels.  The speed/gamma chain uses explicit
// round-to-nearest intrinsics so that it is never FMA-contracted.
#pragma once
#include "vrt_internal.cuh"

struct Sp { double m, q, pmin, m_inv; };

__device__ __forceinline__ long NS(const VrtPatchDev& P, int i, int j) { return (long)P.pitch * (i + 2) + 2 + j; }
// Rectangle::Momentum (Rectangle.hpp:86-88)
__device__ __forceinline__ double momentum(const VrtPatchDev& P, const Sp& sp, double i) {
    return __dadd_rn(sp.pmin, __dmul_rn(P.dp, __dadd_rn(i, (double)P.p_pos)));
}
// Rectangle::Gamma (Rectangle.hpp:181-183)
__device__ __forceinline__ double gamma_(const Sp& sp, double p, double a2) {
    double k = __dmul_rn(__dmul_rn(sp.m_inv, VRT_C_INV), __dmul_rn(sp.m_inv, VRT_C_INV));
    return __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(p, p), a2), k)));
}
__device__ __forceinline__ int finest_index(const VrtPatchDev& P, int i) { return (int)((double)P.rtb * (double)(i + P.x_pos)); }
__device__ __forceinline__ double a_sq(const VrtFields& F, int i) { return F.a_squared[min(max(i, 0), F.N)]; }
// Rectangle::GetEfield (Rectangle.cpp:1055-1067) on the tabulated EMFieldSolver::GetEfield
__device__ __forceinline__ double patch_efield(const VrtPatchDev& P, const VrtFields& F, int i) {
    int j = finest_index(P, i);
    double t = 0.0;
    for (int k = 0; k < P.rtb; k++) t += F.E[j + k + F.epad];
    t *= (1.0 / (double)P.rtb);
    return t;
}
__device__ __forceinline__ double vmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double vmin(double a, double b) { return a < b ? a : b; }

// Rectangle::GetWenoEdgeValueNoMax (Rectangle.cpp:980-1030)
__device__ __forceinline__ double weno(double f1, double f2, double f3, double f4, bool right) {
    double fL = (1.0 / 6) * (-f1 + 5 * f2 + 2 * f3);
    double fR = (1.0 / 6) * (2 * f2 + 5 * f3 - f4);
    double AL = f1 - 2 * f2 + f3, BL = f3 - f1;
    double AR = f2 - 2 * f3 + f4, BR = f4 - f2;
    double bL = 4.0 / 3 * (AL * AL) + 0.5 * AL * BL + 0.25 * (BL * BL);
    double bR = 4.0 / 3 * (AR * AR) - 0.5 * AR * BR + 0.25 * (BR * BR);
    double mm = 1.0e-10;
    double oL = 0.5 / ((mm + bL) * (mm + bL));
    double oR = 0.5 / ((mm + bR) * (mm + bR));
    double wL = oL / (oL + oR), wR = oR / (oL + oR);
    double wL0 = wL * (0.75 + wL * (wL - 0.5));
    double wR0 = wR * (0.75 + wR * (wR - 0.5));
    double W = right ? ((wL0 > wR0) ? wL0 : wR0) : ((wL0 < wR0) ? wL0 : wR0);
    double a = W / (wL0 + wR0);
    return a * fL + (1 - a) * fR;
}


// ---- fused-path arithmetic helpers ---------------------------------------------------------------------------
// reciprocal of a well-scaled operand (no denormal / inf / nan): MUFU.RCP64H seed (relative error e0 < 2^-20) and one
// third-order step x (1 + e + e^2), e = 1 - d x: error e0^3 plus rounding, < 1 ulp, in three fp64 instructions
__device__ __forceinline__ double rcp_scaled(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);
    return fma(x, fma(e, e, e), x);
}
// Rectangle::GetWenoEdgeValueNoMax (Rectangle.cpp:980-1030) with the five divisions folded into one.
//   wL = oL/(oL+oR) = DR/(DL+DR),  D = (1e-10+b)^2;   wL0 = wL(0.75+wL(wL-0.5))  =>  wL0 (DL+DR)^3 = NL
//   NL = DR (0.75 S^2 + DR^2 - 0.5 DR S),  NR likewise with DL;   W/(wL0+wR0) = sel(NL,NR)/(NL+NR)
// DL, DR are first scaled by the power of two that brings S = DL+DR into [1,2) (exact), so S^3 cannot overflow.
// Differs from the reference's evaluation order by a few ulp of the weights (continuous; SURVEY.md H2 allows it).
// common tail: candidates fL, fR and smoothness indicators bL, bR -> face value
// fLmR = fL - fR
__device__ __forceinline__ double weno_fast_tail(double fLmR, double fR, double bL, double bR, bool right) {
    const double mm = 1.0e-10;
    const double DL = (mm + bL) * (mm + bL), DR = (mm + bR) * (mm + bR);
    const double S = DL + DR;
    // sc = 2^(1023 - exponent(S)): S > 0 and finite (D ~ f^4 stays far below 2^1023 for any f the reference itself survives)
    const double sc = __hiloint2double(0x7fe00000 - (__double2hiint(S) & 0x7ff00000), 0);
    const double dl = DL * sc, dr = DR * sc, ss = dl + dr;
    const double h = 0.75 * ss;
    const double NL = dr * fma(ss, fma(-0.5, dr, h), dr * dr);
    const double NR = dl * fma(ss, fma(-0.5, dl, h), dl * dl);
    // right: the larger weight, left: the smaller one; on a tie both candidates are equal
    const bool pickL = (NL > NR) == right;
    const double a = (pickL ? NL : NR) * rcp_scaled(NL + NR);
    return fma(a, fLmR, fR);                     // a fL + (1 - a) fR
}
__device__ __forceinline__ double weno_fast(double f1, double f2, double f3, double f4, bool right) {
    const double fR = (1.0 / 6) * (2 * f2 + 5 * f3 - f4);
    const double AL = f1 - 2 * f2 + f3, BL = f3 - f1;
    const double AR = f2 - 2 * f3 + f4, BR = f4 - f2;
    const double bL = 4.0 / 3 * (AL * AL) + 0.5 * AL * BL + 0.25 * (BL * BL);
    const double bR = 4.0 / 3 * (AR * AR) - 0.5 * AR * BR + 0.25 * (BR * BR);
    return weno_fast_tail((1.0 / 6) * (AR - AL), fR, bL, bR, right);      // fL - fR = (f4 - 3 f3 + 3 f2 - f1) / 6 = (AR - AL) / 6
}
// The same for a stencil that slides by one cell per call (the x direction of the marching kernel): the right candidate's
// second difference A and first difference B are the left candidate's of the next face, and the two smoothness indicators
// differ only in the sign of the cross term, so each call forms one (A, B) pair, bR = t1 - t2, and hands bL = t1 + t2 of the
// next face on through `bL_next` (in: this face's bL from the previous call).
__device__ __forceinline__ double weno_fast_sliding(double f1, double f2, double f3, double f4, bool right, double& bL_next) {
    const double fL = (1.0 / 6) * (-f1 + 5 * f2 + 2 * f3);
    const double fR = (1.0 / 6) * (2 * f2 + 5 * f3 - f4);
    const double AR = f2 - 2 * f3 + f4, BR = f4 - f2;
    const double t1 = fma(0.25 * BR, BR, 4.0 / 3 * (AR * AR)), t2 = (0.5 * AR) * BR;
    const double bL = bL_next, bR = t1 - t2;
    bL_next = t1 + t2;
    return weno_fast_tail(fL - fR, fR, bL, bR, right);
}
