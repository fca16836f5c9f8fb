/* TEST INFRASTRUCTURE — CPU oracle.  Plain-C restatement of the reference's 1D1P Vlasov advance
 * (Libbum/Veritas, /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (veritas_b200/) never does.
 *
 * Parity pinning: the reference has no tests or golden vectors of its own.  This port
 * is pinned against the reference ITSELF: oracle/_ref/ref_harness (the unmodified reference sources
 * compiled by oracle/Makefile) dumps full-precision state, and tests/test_oracle_vs_reference.py
 * checks this file against those dumps (committed as tests/golden/ *.bin fixtures with the
 * generating script tests/golden/make_golden.py).  The third-party arithmetic on the path (LAPACK
 * dgetrf/dgetrs, MKL vdLn — un-pinned versions) is restated as a plain partial-pivot
 * LU and libm log; for those the parity is "unpinned" against MKL and pinned against the OpenBLAS
 * 0.3.15 build of the reference used here.
 *
 * Build: gcc -O2 -ffp-contract=off (x86-64 baseline, no FMA: SURVEY.md H2).
 *
 * Layout: p is the fast index; a patch array has (n_x+4)*(n_p+4) doubles, cell (i,j) at
 * NS(i,j) = (n_p+4)*(i+2)+2+j (Rectangle.hpp:105-108).  The reference's AoS f[3*idx+state] and
 * F?H[6*idx+slot] are held here as separate planes (f0,f1,f2; slot-major flux history).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS0_INV 1.1294e+11      /* veritas.hpp:22 */
#define MU_INV 795774.715482     /* veritas.hpp:25 */
#define CS 299792458.0           /* veritas.hpp:26 */
#define C_INV 3.33564095e-9      /* veritas.hpp:27 */

typedef struct {
    int n_x, n_p, x_pos, p_pos, up, down, left, right, rtb, pad_;
    double dx, dp, pmin, m, q;
    double *f0, *f1, *f2, *fx, *fp, *ex, *ep;
    double *FxH, *FpH;   /* 6 planes each, slot-major */
    double *FxL, *FpL;   /* slot 0 only is ever read (quirk Q1); one plane each, written at step 0 */
    double *FxLS, *FpLS, *FxDS, *FpDS, *Rp, *Rm, *Cx, *Cp;
} vo_patch;

typedef struct {
    int x_size, n_prepad, n_postpad, pad_;
    double dx;                      /* finest dx */
    double *By, *Bz, *Ey, *Ez, *Ay, *Az;   /* 8 slots x M, slot-major: Index(i,s)=s*M+i (EMSolver.hpp:51-53) */
    double *a_squared;              /* x_size+1 */
    double *PHI, *charge, *J, *neutral;
    double Ex0;
} vo_fields;

static inline long NS(const vo_patch* P, int i, int j) { return (long)(P->n_p + 4) * (i + 2) + 2 + j; }
static inline double vmax(double a, double b) { return a > b ? a : b; }   /* Rectangle.hpp:110-117 */
static inline double vmin(double a, double b) { return a < b ? a : b; }   /* Rectangle.hpp:119-126 */

/* Rectangle::Momentum (Rectangle.hpp:86-88): lower p-face of cell i */
static inline double momentum(const vo_patch* P, double i) { return P->pmin + P->dp * (i + P->p_pos); }
/* Rectangle::Gamma (Rectangle.hpp:181-183) */
static inline double gamma_(const vo_patch* P, double p, double a2) {
    double m_inv = 1 / P->m;
    return sqrt(1.0 + ((p * p) + a2) * ((m_inv * C_INV) * (m_inv * C_INV)));
}
/* Rectangle::GetFinestIndex (Rectangle.hpp:82-84) */
static inline int finest_index(const vo_patch* P, int i) { return (int)((double)P->rtb * (i + P->x_pos)); }
/* EMFieldSolver::GetASquared (EMSolver.cpp:133-135) */
static inline double a_sq(const vo_fields* F, int i) {
    int k = i < 0 ? 0 : i; if (k > F->x_size) k = F->x_size;
    return F->a_squared[k];
}
/* EMFieldSolver::GetEfield (EMSolver.cpp:137-154) */
double vo_em_efield(const vo_fields* F, int i) {
    int N = F->x_size;
    int ip1 = i + 1, im1 = i - 1, ip2 = i + 2, im2 = i - 2;
    ip1 = ip1 > -1 ? ip1 : ip1 + N; im1 = im1 > -1 ? im1 : im1 + N;
    ip2 = ip2 > -1 ? ip2 : ip2 + N; im2 = im2 > -1 ? im2 : im2 + N;
    ip1 = ip1 < N ? ip1 : ip1 - N; im1 = im1 < N ? im1 : im1 - N;
    ip2 = ip2 < N ? ip2 : ip2 - N; im2 = im2 < N ? im2 : im2 - N;
    double fieldCoef = 1.0 / (12 * F->dx);   /* EMSolver.cpp:88 */
    return -fieldCoef * (8 * (F->PHI[ip1] - F->PHI[im1]) - F->PHI[ip2] + F->PHI[im2]) + F->Ex0;
}
/* Rectangle::GetEfield (Rectangle.cpp:1055-1067) */
static double patch_efield(const vo_patch* P, const vo_fields* F, int i) {
    int j = finest_index(P, i);
    double t = 0.0;
    for (int k = 0; k < P->rtb; k++) t += vo_em_efield(F, j + k);
    t *= (1.0 / (double)P->rtb);
    return t;
}
/* Rectangle::GetWenoEdgeValueNoMax (Rectangle.cpp:980-1030); the scalar twin (943-977) is bit-identical */
double vo_weno(double f1, double f2, double f3, double f4, int right) {
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
    double W;
    if (right) W = (wL0 > wR0) ? wL0 : wR0; else W = (wL0 < wR0) ? wL0 : wR0;
    double a = W / (wL0 + wR0);
    return a * fL + (1 - a) * fR;
}

/* RK tableau rows (Rectangle.cpp:1397-1498, EMSolver.cpp:210-312), literal * timestep */
static const double RK_A[6][6] = {
    {0.5, 0, 0, 0, 0, 0},
    {0.221776, 0.110224, 0, 0, 0, 0},
    {-0.04884659515311857, -0.17772065232640102, 0.8465672474795197, 0, 0, 0},
    {-0.15541685842491548, -0.3567050098221991, 1.0587258798684427, 0.30339598837867193, 0, 0},
    {0.2014243506726763, 0.008742057842904185, 0.15993995707168115, 0.4038290605220775, 0.22606457389066084, 0},
    {0.15791629516167136, 0.0, 0.18675894052400077, 0.6805652953093346, -0.27524053099500667, 0.25}};

/* Rectangle::FCTTimeStep (Rectangle.cpp:1255-1623) for a patch without interior level boundaries. */
void vo_fct_substep(vo_patch* P, const vo_fields* F, double timestep, int step, int subStep) {
    const int nx = P->n_x, np = P->n_p;
    const int xm = P->left ? 1 : 0, xp = P->right ? nx : nx + 1;
    const int pp = P->up ? np : np + 1, pm = P->down ? 1 : 0;
    const long npad = (long)(nx + 4) * (np + 4);
    if (subStep == 0) {
        const double q = P->q, w3 = 1 / 48.0, dx_inv = 1 / P->dx, dp_inv = 1 / P->dp, cc = CS * CS * P->m;
        #pragma omp parallel
        {
        #pragma omp for schedule(static)
        for (int i = -1; i < nx + 2; i++) {
            double as = q * q * a_sq(F, finest_index(P, i));
            for (int j = -1; j < np + 1; j++)
                P->ex[NS(P, i, j)] = dp_inv * cc * (gamma_(P, momentum(P, j + 1), as) - gamma_(P, momentum(P, j), as));
        }
        #pragma omp for schedule(static)
        for (int i = 0; i < nx + 1; i++)
            for (int j = -1; j < np + 1; j++)
                P->fx[NS(P, i, j)] = vo_weno(P->f1[NS(P, i - 2, j)], P->f1[NS(P, i - 1, j)], P->f1[NS(P, i, j)],
                                             P->f1[NS(P, i + 1, j)], P->ex[NS(P, i, j)] > 0.0);
        #pragma omp for schedule(static)
        for (int i = -1; i < nx + 1; i++) {
            double as_1 = q * q * a_sq(F, finest_index(P, i));
            double as_2 = q * q * a_sq(F, finest_index(P, i + 1));
            double Em = q * patch_efield(P, F, i);
            for (int j = -1; j < np + 2; j++) {
                double mom = momentum(P, j);
                P->ep[NS(P, i, j)] = Em - cc * dx_inv * (gamma_(P, mom, as_2) - gamma_(P, mom, as_1));
            }
        }
        #pragma omp for schedule(static)
        for (int i = -1; i < nx + 1; i++)
            for (int j = 0; j < np + 1; j++)
                P->fp[NS(P, i, j)] = vo_weno(P->f1[NS(P, i, j - 2)], P->f1[NS(P, i, j - 1)], P->f1[NS(P, i, j)],
                                             P->f1[NS(P, i, j + 1)], P->ep[NS(P, i, j)] > 0.0);
        double* FxHs = P->FxH + step * npad; double* FpHs = P->FpH + step * npad;
        #pragma omp for schedule(static)
        for (int i = 0; i < nx + 1; i++)
            for (int j = -1; j < np + 1; j++) {
                long c = NS(P, i, j);
                double am = P->ex[c], ap1 = P->ex[c + 1], am1 = P->ex[c - 1];
                double fm = P->fx[c], fp1 = P->fx[c + 1], fm1 = P->fx[c - 1];
                FxHs[c] = dx_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
                if (step == 0) P->FxL[c] = dx_inv * ((am > 0.0 ? P->f1[NS(P, i - 1, j)] : P->f1[c]) * am);
            }
        #pragma omp for schedule(static)
        for (int i = -1; i < nx + 1; i++)
            for (int j = 0; j < np + 1; j++) {
                long c = NS(P, i, j), cp = NS(P, i + 1, j), cm = NS(P, i - 1, j);
                double am = P->ep[c], ap1 = P->ep[cp], am1 = P->ep[cm];
                double fm = P->fp[c], fp1 = P->fp[cp], fm1 = P->fp[cm];
                FpHs[c] = dp_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
                if (step == 0) P->FpL[c] = dp_inv * ((am > 0.0 ? P->f1[c - 1] : P->f1[c]) * am);
            }
        }
        /* RK combination over the whole padded array (Rectangle.cpp:1396-1517) */
        double a[6], aSum = 0.0;
        for (int k = 0; k <= step; k++) { a[k] = RK_A[step][k] * timestep; aSum = (k == 0) ? a[0] : aSum + a[k]; }
        #pragma omp parallel for schedule(static)
        for (long c = 0; c < npad; c++) {
            P->FxLS[c] = aSum * P->FxL[c];
            P->FpLS[c] = aSum * P->FpL[c];
            double sx = a[0] * P->FxH[c], sp = a[0] * P->FpH[c];
            for (int k = 1; k <= step; k++) { sx = sx + a[k] * P->FxH[k * npad + c]; sp = sp + a[k] * P->FpH[k * npad + c]; }
            P->FxDS[c] = sx - P->FxLS[c];
            P->FpDS[c] = sp - P->FpLS[c];
        }
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < np; j++) P->f2[NS(P, i, j)] = P->f0[NS(P, i, j)];
        /* serial scatter, reference order (Rectangle.cpp:1526-1534; quirks Q11, Q14) */
        for (int i = xm; i < xp; i++)
            for (int j = pm; j < pp; j++) {
                long c = NS(P, i, j);
                P->f2[c] += P->FxLS[c];
                P->f2[NS(P, i - 1, j)] -= P->FxLS[c];
                P->f2[c] += P->FpLS[c];
                P->f2[c - 1] -= P->FpLS[c];
            }
    } else if (subStep == 1) {
        #pragma omp parallel for schedule(static)
        for (int i = -1; i < nx + 1; i++)
            for (int j = -1; j < np + 1; j++) {
                long c = NS(P, i, j), cxp = NS(P, i + 1, j), cxm = NS(P, i - 1, j);
                double Pp = vmax(0.0, P->FxDS[c]) - vmin(0.0, P->FxDS[cxp]) + vmax(0.0, P->FpDS[c]) - vmin(0.0, P->FpDS[c + 1]);
                double Pm = vmax(0.0, P->FxDS[cxp]) - vmin(0.0, P->FxDS[c]) + vmax(0.0, P->FpDS[c + 1]) - vmin(0.0, P->FpDS[c]);
                double w1a = vmax(P->f0[c], P->f2[c]), w2a = vmax(P->f0[cxp], P->f2[cxp]), w3a = vmax(P->f0[cxm], P->f2[cxm]);
                double w4a = vmax(P->f0[c + 1], P->f2[c + 1]), w5a = vmax(P->f0[c - 1], P->f2[c - 1]);
                double wMax = vmax(w1a, vmax(w2a, vmax(w3a, vmax(w4a, w5a))));
                double w1i = vmin(P->f0[c], P->f2[c]), w2i = vmin(P->f0[cxp], P->f2[cxp]), w3i = vmin(P->f0[cxm], P->f2[cxm]);
                double w4i = vmin(P->f0[c + 1], P->f2[c + 1]), w5i = vmin(P->f0[c - 1], P->f2[c - 1]);
                double wMin = vmin(w1i, vmin(w2i, vmin(w3i, vmin(w4i, w5i))));
                double Qm = -wMin + P->f2[c], Qp = wMax - P->f2[c];
                P->Rp[c] = Pp > 0.0 ? vmin(1.0, Qp / Pp) : 0.0;
                P->Rm[c] = Pm > 0.0 ? vmin(1.0, Qm / Pm) : 0.0;
            }
        for (long c = 0; c < npad; c++) { P->Cp[c] = 1.0; P->Cx[c] = 1.0; }
        #pragma omp parallel for schedule(static)
        for (int i = 1; i < nx; i++)
            for (int j = 0; j < np; j++) {
                long c = NS(P, i, j), ci = NS(P, i - 1, j), cj = c - 1;
                P->Cx[c] = P->FxDS[c] > 0.0 ? vmin(P->Rp[c], P->Rm[ci]) : vmin(P->Rp[ci], P->Rm[c]);
                P->Cp[c] = P->FpDS[c] > 0.0 ? vmin(P->Rp[c], P->Rm[cj]) : vmin(P->Rp[cj], P->Rm[c]);
            }
    } else if (subStep == 2) {
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < np; j++) P->f1[NS(P, i, j)] = P->f2[NS(P, i, j)];
        for (int i = xm; i < xp; i++)
            for (int j = pm; j < pp; j++) {
                long c = NS(P, i, j);
                P->f1[c] += P->Cx[c] * P->FxDS[c];
                P->f1[NS(P, i - 1, j)] -= P->Cx[c] * P->FxDS[c];
                P->f1[c] += P->Cp[c] * P->FpDS[c];
                P->f1[c - 1] -= P->Cp[c] * P->FpDS[c];
            }
    } else if (subStep == 3) {
        memcpy(P->f0, P->f1, sizeof(double) * npad);
    }
}

/* Ghost fill of state `val` (1 or 2) for a patch whose every neighbour is the physical boundary:
 * UpdateSameLevelBoundaries + UpdateCornerPoints (Rectangle.cpp:562-614, 1130-1214) with
 * BoundaryCondition::GetValueFromSameLevel == 0.0 (BoundaryCondition.cpp:6-8). */
void vo_fill_domain_ghosts(vo_patch* P, int val) {
    double* f = val == 2 ? P->f2 : (val == 1 ? P->f1 : P->f0);
    for (int i = -2; i < P->n_x + 2; i++)
        for (int j = -2; j < P->n_p + 2; j++)
            if (i < 0 || i >= P->n_x || j < 0 || j >= P->n_p) f[NS(P, i, j)] = 0.0;
}

/* Mesh::Advance for one single-patch level (Mesh.cpp:64-89) */
void vo_mesh_advance_single(vo_patch* P, const vo_fields* F, double dt, int step) {
    vo_fct_substep(P, F, dt, step, 0);
    vo_fill_domain_ghosts(P, 2);
    vo_fct_substep(P, F, dt, step, 1);
    /* PushBoundaryC: every neighbour is the BoundaryCondition object -> zero-trip loops (quirk Q8) */
    vo_fct_substep(P, F, dt, step, 2);
    vo_fill_domain_ghosts(P, 1);
    if (step == 5) vo_fct_substep(P, F, dt, step, 3);
}

/* EMFieldSolver::GetCellAverageASquared (EMSolver.hpp:56-63) */
static inline double cell_a_sq(const vo_fields* F, int i) {
    int M = F->x_size + F->n_prepad + F->n_postpad;
    i += F->n_prepad; i = i > -1 ? i : 0; i = i < M ? i : M - 1;
    double ay = F->Ay[M + i], az = F->Az[M + i];
    return (ay * ay) + (az * az);
}

static const double IM[12] = {0.104166666666667, -0.708333333333334, 0.708333333333334, -0.104166666666667,
                              0.117647058823529, 0.029411764705882,  0.029411764705882, 0.117647058823529,
                              -0.083333333333333, 0.166666666666667, -0.166666666666667, 0.083333333333334};
/* sub-cell moments of the cubic (Rectangle.cpp:94-111) */
static void interp_coefs(int n, double* c) {
    for (int i = 0; i < n; i++) {
        double tl = -0.5 + i / (double)n, tr = -0.5 + (i + 1.0) / (double)n;
        c[3 * i] = (tl + tr) * 0.5;
        c[3 * i + 1] = (tl * tl + tl * tr + tr * tr) / 3.0 - (1.0 / 12);
        c[3 * i + 2] = (tl * tl * tl + tl * tl * tr + tl * tr * tr + tr * tr * tr) * 0.25;
    }
}
/* Rectangle::GetInterpolantsREL / REF (Rectangle.cpp:121-155) */
static void interpolants(const double* coefs, int n, double f1, double f2, double f3, double f4, double f5, double* out) {
    f5 -= f3; f4 -= f3; f2 -= f3; f1 -= f3;
    double a1 = IM[0] * f1 + IM[1] * f2 + IM[2] * f4 + IM[3] * f5;
    double a2 = IM[4] * f1 + IM[5] * f2 + IM[6] * f4 + IM[7] * f5;
    double a3 = IM[8] * f1 + IM[9] * f2 + IM[10] * f4 + IM[11] * f5;
    for (int i = 0; i < n; i++) out[i] = coefs[3 * i] * a1 + coefs[3 * i + 1] * a2 + coefs[3 * i + 2] * a3 + f3;
}

/* Rectangle::CalculateRhoAndJ, USINGMKL branch (Rectangle.cpp:157-282, quirk Q7).  nested may be NULL.
 * chargeR/currentR have n_x*rtb entries. */
void vo_patch_moments(const vo_patch* P, const vo_fields* F, const unsigned char* nested, double* chargeR, double* currentR) {
    const double q = P->q, mass = P->m, m_inv = 1 / P->m;
    const double c1 = m_inv * C_INV, c2 = 1 / c1, c3 = 1 / 48.0;
    const int rtb = P->rtb, nx = P->n_x, np = P->n_p;
    double* coefs = (double*)malloc(sizeof(double) * 3 * rtb);
    interp_coefs(rtb, coefs);
    const double rel = 1.0 / (double)rtb;
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; i++) {
        double* buf = (double*)malloc(sizeof(double) * 3 * rtb);
        double *tm1 = buf, *t0 = buf + rtb, *tp1 = buf + 2 * rtb;
        for (int k = 0; k < rtb; k++) { chargeR[i * rtb + k] = 0.0; currentR[i * rtb + k] = 0.0; }
        for (int r = 0; r < 3; r++) {   /* rows j = -1, 0, 1 */
            int j = r - 1; double* t = r == 0 ? tm1 : (r == 1 ? t0 : tp1);
            interpolants(coefs, rtb, P->f1[NS(P, i - 2, j)], P->f1[NS(P, i - 1, j)], P->f1[NS(P, i, j)], P->f1[NS(P, i + 1, j)], P->f1[NS(P, i + 2, j)], t);
            double s = 0.0; for (int k = 0; k < rtb; k++) s += t[k];
            double cor = P->f1[NS(P, i, j)] - rel * s;
            for (int k = 0; k < rtb; k++) t[k] += cor;
        }
        for (int j = 0; j < np; j++) {
            if (!(nested && nested[NS(P, i, j)])) {
                for (int k = 0; k < rtb; k++) {
                    double a2 = cell_a_sq(F, (i + P->x_pos) * rtb + k);
                    chargeR[i * rtb + k] += t0[k];
                    double gm1_ = gamma_(P, momentum(P, j - 1), q * q * a2), g0 = gamma_(P, momentum(P, j), q * q * a2);
                    double g1 = gamma_(P, momentum(P, j + 1), q * q * a2), g2 = gamma_(P, momentum(P, j + 2), q * q * a2);
                    double l0 = (g1 + c1 * momentum(P, j + 1)) / (g0 + c1 * momentum(P, j));
                    double lm = (g0 + c1 * momentum(P, j)) / (gm1_ + c1 * momentum(P, j - 1));
                    double lp = (g2 + c1 * momentum(P, j + 2)) / (g1 + c1 * momentum(P, j + 1));
                    double gm = c2 * log(l0), gmm = c2 * log(lm), gmp = c2 * log(lp);
                    currentR[i * rtb + k] += t0[k] * gm + c3 * (gmp - gmm) * (tp1[k] - tm1[k]);
                }
            }
            double* t = tm1; tm1 = t0; t0 = tp1; tp1 = t;
            interpolants(coefs, rtb, P->f1[NS(P, i - 2, j + 2)], P->f1[NS(P, i - 1, j + 2)], P->f1[NS(P, i, j + 2)], P->f1[NS(P, i + 1, j + 2)], P->f1[NS(P, i + 2, j + 2)], tp1);
            double s = 0.0; for (int k = 0; k < rtb; k++) s += tp1[k];
            double cor = P->f1[NS(P, i, j + 2)] - rel * s;
            for (int k = 0; k < rtb; k++) tp1[k] += cor;
        }
        for (int k = 0; k < rtb; k++) { chargeR[i * rtb + k] *= P->dp * q; currentR[i * rtb + k] *= -q * q / mass; }
        free(buf);
    }
    free(coefs);
}

/* Dense periodic 4th-order Poisson matrix of EMFieldSolver::EMFieldSolver (EMSolver.cpp:28-67), built in the
 * column-major sense LAPACK sees, factored by unblocked partial-pivot LU (the algorithm of LAPACK dgetf2;
 * dgetrf/dgetrs themselves live in the un-vendored MKL/LAPACK dependency). */
typedef struct { int n; double* lu; int* piv; } vo_poisson;
vo_poisson* vo_poisson_create(int n) {
    vo_poisson* S = (vo_poisson*)malloc(sizeof(vo_poisson));
    S->n = n; S->lu = (double*)calloc((size_t)n * n, sizeof(double)); S->piv = (int*)malloc(sizeof(int) * n);
    double w1 = -16.0 / 30.0 / 0.4, w2 = 1.0 / 30.0 / 0.4, w3 = 1 / 0.4;
    double* pM = S->lu;   /* pM[c*n + r]: LAPACK column c, row r == the reference's pM[i*x_size + k] with i=c,k=r */
    for (int i = 2; i < n - 2; i++) { pM[i * n + i] = w3; pM[i * n + i + 1] = w1; pM[i * n + i - 1] = w1; pM[i * n + i + 2] = w2; pM[i * n + i - 2] = w2; }
    pM[0] = 1.0;
    pM[n + 1] = w3; pM[n + 2] = w1; pM[n + 0] = w1; pM[n + 3] = w2; pM[2 * n - 1] = w2;
    pM[(n - 2) * n + n - 2] = w3; pM[(n - 2) * n + n - 1] = w1; pM[(n - 2) * n + n - 3] = w1; pM[(n - 2) * n + 0] = w2; pM[(n - 2) * n + n - 4] = w2;
    pM[(n - 1) * n + (n - 1)] = w3; pM[(n - 1) * n + 0] = w1; pM[(n - 1) * n + (n - 2)] = w1; pM[(n - 1) * n + 1] = w2; pM[(n - 1) * n + (n - 3)] = w2;
    /* right-looking LU with partial pivoting, column-major; exploits the band to skip zero multipliers */
    for (int k = 0; k < n; k++) {
        int p = k; double best = fabs(pM[k * n + k]);
        for (int r = k + 1; r < n; r++) if (fabs(pM[k * n + r]) > best) { best = fabs(pM[k * n + r]); p = r; }
        S->piv[k] = p;
        if (p != k) for (int c = 0; c < n; c++) { double t = pM[c * n + k]; pM[c * n + k] = pM[c * n + p]; pM[c * n + p] = t; }
        double d = pM[k * n + k];
        for (int r = k + 1; r < n; r++) if (pM[k * n + r] != 0.0) pM[k * n + r] /= d;
        #pragma omp parallel for schedule(static)
        for (int c = k + 1; c < n; c++) {
            double u = pM[c * n + k];
            if (u != 0.0) for (int r = k + 1; r < n; r++) { double l = pM[k * n + r]; if (l != 0.0) pM[c * n + r] -= l * u; }
        }
    }
    return S;
}
void vo_poisson_destroy(vo_poisson* S) { free(S->lu); free(S->piv); free(S); }
static void poisson_solve(const vo_poisson* S, double* b) {
    int n = S->n; const double* A = S->lu;
    for (int k = 0; k < n; k++) { int p = S->piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int k = 0; k < n; k++) { double v = b[k]; if (v != 0.0) for (int r = k + 1; r < n; r++) b[r] -= A[k * n + r] * v; }
    for (int k = n - 1; k >= 0; k--) { b[k] /= A[k * n + k]; double v = b[k]; if (v != 0.0) for (int r = 0; r < k; r++) b[r] -= A[k * n + r] * v; }
}
/* EMFieldSolver::UpdatePotential (EMSolver.cpp:156-192) */
void vo_update_potential(const vo_poisson* S, vo_fields* F) {
    int N = F->x_size;
    double te = EPS0_INV;
    for (int i = 0; i < N; i++) F->PHI[i] = te * (F->charge[i] + F->neutral[i]);
    double w = F->dx * F->dx;   /* std::pow(dx, 2.0) */
    for (int i = 0; i < N; i++) F->PHI[i] *= w;
    poisson_solve(S, F->PHI);
    F->Ex0 += -(vo_em_efield(F, -1) + vo_em_efield(F, 0)) * 0.5;
}
/* Ex0 update alone, for runs that inject PHI from a reference dump */
void vo_update_ex0(vo_fields* F) { F->Ex0 += -(vo_em_efield(F, -1) + vo_em_efield(F, 0)) * 0.5; }

/* EMFieldSolver::RGKCalculateRHS (EMSolver.cpp:479-553); by0/bz0 = Settings::GetBY/GetBZ(0,time) */
void vo_field_rhs(vo_fields* F, int step, double by0, double bz0) {
    const int N = F->x_size, pre = F->n_prepad, post = F->n_postpad, M = N + pre + post;
    double *By = F->By, *Bz = F->Bz, *Ey = F->Ey, *Ez = F->Ez, *Ay = F->Ay, *Az = F->Az;
    const long s2 = (long)(step + 2) * M, s1 = M;
    for (int i = 0; i < M; i++) { By[s2 + i] = 0; Bz[s2 + i] = 0; Ey[s2 + i] = 0; Ez[s2 + i] = 0; Ay[s2 + i] = 0; Az[s2 + i] = 0; }
    const double dx_inv = 1 / F->dx;
    {
        int i = 0;
        By[s2 + i] = dx_inv * (Ez[s1 + i + 1] - Ez[s1 + i]);
        Bz[s2 + i] = -dx_inv * (Ey[s1 + i + 1] - Ey[s1 + i]);
        Ey[s2 + i] = -dx_inv * (Bz[s1 + i] - bz0) * EPS0_INV * MU_INV;
        Ez[s2 + i] = dx_inv * (By[s1 + i] - by0) * EPS0_INV * MU_INV;
        Ay[s2 + i] = -Ey[s1 + i];
        Az[s2 + i] = -Ez[s1 + i];
    }
    for (int i = 1; i < pre; i++) {
        By[s2 + i] = dx_inv * (Ez[s1 + i + 1] - Ez[s1 + i]);
        Bz[s2 + i] = -dx_inv * (Ey[s1 + i + 1] - Ey[s1 + i]);
        Ey[s2 + i] = -dx_inv * (Bz[s1 + i] - Bz[s1 + i - 1]) * EPS0_INV * MU_INV;
        Ez[s2 + i] = dx_inv * (By[s1 + i] - By[s1 + i - 1]) * EPS0_INV * MU_INV;
        Ay[s2 + i] = -Ey[s1 + i];
        Az[s2 + i] = -Ez[s1 + i];
    }
    const double c1 = -1.0 / 24, c2 = 9.0 / 8.0;
    for (int i = pre; i < pre + N; i++) {
        By[s2 + i] = dx_inv * (c1 * (Ez[s1 + i + 2] - Ez[s1 + i - 1]) + c2 * (Ez[s1 + i + 1] - Ez[s1 + i]));
        Bz[s2 + i] = -dx_inv * (c1 * (Ey[s1 + i + 2] - Ey[s1 + i - 1]) + c2 * (Ey[s1 + i + 1] - Ey[s1 + i]));
        Ey[s2 + i] = -dx_inv * (c1 * (Bz[s1 + i + 1] - Bz[s1 + i - 2]) + c2 * (Bz[s1 + i] - Bz[s1 + i - 1])) * EPS0_INV * MU_INV - EPS0_INV * F->J[i - pre] * Ay[s1 + i];
        Ez[s2 + i] = dx_inv * (c1 * (By[s1 + i + 1] - By[s1 + i - 2]) + c2 * (By[s1 + i] - By[s1 + i - 1])) * EPS0_INV * MU_INV - EPS0_INV * F->J[i - pre] * Az[s1 + i];
        Ay[s2 + i] = -Ey[s1 + i];
        Az[s2 + i] = -Ez[s1 + i];
    }
    for (int i = pre + N; i < pre + N + post - 2; i++) {
        By[s2 + i] = dx_inv * (Ez[s1 + i + 1] - Ez[s1 + i]);
        Bz[s2 + i] = -dx_inv * (Ey[s1 + i + 1] - Ey[s1 + i]);
        Ey[s2 + i] = -dx_inv * (Bz[s1 + i] - Bz[s1 + i - 1]) * EPS0_INV * MU_INV;
        Ez[s2 + i] = dx_inv * (By[s1 + i] - By[s1 + i - 1]) * EPS0_INV * MU_INV;
        Ay[s2 + i] = -Ey[s1 + i];
        Az[s2 + i] = -Ez[s1 + i];
    }
}
/* EMFieldSolver::RGKUpdateIntermediateSolution (EMSolver.cpp:204-338) */
void vo_field_update(vo_fields* F, int step, double timestep) {
    const int M = F->x_size + F->n_prepad + F->n_postpad;
    double* Y[6] = {F->By, F->Bz, F->Ey, F->Ez, F->Ay, F->Az};
    if (step < 5) {
        double a[5];
        for (int k = 0; k <= step; k++) a[k] = RK_A[step][k] * timestep;
        for (int v = 0; v < 6; v++) {
            double* y = Y[v];
            for (int i = 0; i < M; i++) {
                double s = y[i] + a[0] * y[2 * M + i];
                for (int k = 1; k <= step; k++) s = s + a[k] * y[(long)(2 + k) * M + i];
                y[M + i] = s;
            }
        }
    } else {
        double b[6];
        for (int k = 0; k < 6; k++) b[k] = RK_A[5][k] * timestep - (k < 5 ? RK_A[4][k] * timestep : 0);
        for (int v = 0; v < 6; v++) {
            double* y = Y[v];
            for (int i = 0; i < M; i++) {
                double s = y[M + i] + b[0] * y[2 * M + i];
                for (int k = 1; k < 6; k++) s = s + b[k] * y[(long)(2 + k) * M + i];
                y[i] = s; y[M + i] = s;
            }
        }
    }
}
static double weno_unbiased(double f1, double f2, double f3, double f4) {   /* EMSolver.cpp:565-585 */
    double fL = (1.0 / 6) * (-f1 + 5 * f2 + 2 * f3), fR = (1.0 / 6) * (2 * f2 + 5 * f3 - f4);
    double AL = f1 - 2 * f2 + f3, BL = f3 - f1, AR = f2 - 2 * f3 + f4, BR = f4 - f2;
    double bL = 4.0 / 3 * (AL * AL) + 0.5 * AL * BL + 0.25 * (BL * BL);
    double bR = 4.0 / 3 * (AR * AR) - 0.5 * AR * BR + 0.25 * (BR * BR);
    double mm = 1.0e-10;
    double oL = 0.5 / ((mm + bL) * (mm + bL)), oR = 0.5 / ((mm + bR) * (mm + bR));
    double wL = oL / (oL + oR), wR = oR / (oL + oR);
    return wL * fL + wR * fR;
}
/* EMFieldSolver::InterpolateToFaces (EMSolver.cpp:555-619) */
void vo_field_faces(vo_fields* F) {
    const int N = F->x_size, pre = F->n_prepad, M = N + pre + F->n_postpad;
    const double *Ay = F->Ay + M, *Az = F->Az + M;
    for (int i = 0; i < N; i++) {
        double ay = weno_unbiased(Ay[pre + i - 2], Ay[pre + i - 1], Ay[pre + i], Ay[pre + i + 1]);
        double az = weno_unbiased(Az[pre + i - 2], Az[pre + i - 1], Az[pre + i], Az[pre + i + 1]);
        F->a_squared[i] = (ay * ay) + (az * az);
    }
}
void vo_field_stage(vo_fields* F, int step, double dt, double by0, double bz0) {   /* RGKStep, EMSolver.cpp:194-202 */
    vo_field_rhs(F, step, by0, bz0); vo_field_update(F, step, dt); vo_field_faces(F);
}
/* EMFieldSolver::EstimateCFLBound (EMSolver.cpp:631-664), including the un-offset indexing (quirk Q3) */
double vo_cfl_bound(const vo_fields* F, int n_species, const double* m, const double* q, const double* dp_finest) {
    double dps[8], dpsMax = 0.0;
    for (int i = 0; i < n_species; i++) { dps[i] = 1 / dp_finest[i]; dpsMax = fmax(dpsMax, fabs(q[i]) * dps[i]); }
    double pc = 0.0;
    for (int i = 0; i < F->x_size; i++) {
        double Azv = F->Az[i], Ayv = F->Ay[i], As = Ayv * Ayv + Azv * Azv, t = 0.0;
        for (int j = 0; j < n_species; j++) t = fmax(t, fabs(q[j]) / m[j] / sqrt(1 + As / ((m[j] * CS) * (m[j] * CS))) * dps[j]);
        pc = fmax(pc, t * fabs(Ayv * F->Bz[i] - Azv * F->By[i]) + dpsMax * fabs(vo_em_efield(F, i)));
    }
    return 1.0 / fmax((CS / F->dx + pc), 1e-40);
}
/* Settings::UpdateTime (Settings.cpp:166-179) */
double vo_update_time(double time, int step, double dt) {
    if (step == 1) time += (0.5 * dt);
    else if (step == 2) time += (0.332 - 0.5) * dt;
    else if (step == 3) time += (0.62 - 0.332) * dt;
    else if (step == 4) time += (0.85 - 0.62) * dt;
    else if (step == 5) time += (1.0 - 0.85) * dt;
    return time;
}
/* EMFieldSolver::AssembleRhoAndJ for single-patch, single-level species (EMSolver.cpp:104-122, Level.cpp:19-62) */
void vo_assemble_single(int n_species, vo_patch** P, vo_fields* F, double** charges, double* scratchJ) {
    int N = F->x_size;
    for (int i = 0; i < N; i++) { F->charge[i] = 0.0; F->J[i] = 0.0; }
    for (int s = 0; s < n_species; s++) {
        vo_patch_moments(P[s], F, NULL, charges[s], scratchJ);
        for (int i = 0; i < N; i++) {
            double cl = 0.0 + charges[s][i], jl = 0.0 + scratchJ[i];   /* chargeL / currentL accumulation */
            charges[s][i] = 0.0 + cl;
            F->J[i] += jl;
        }
    }
    for (int s = 0; s < n_species; s++) for (int i = 0; i < N; i++) F->charge[i] += charges[s][i];
}
