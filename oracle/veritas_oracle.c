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

/* Rectangle::FCTTimeStep (Rectangle.cpp:1255-1623).  M/pidx (may be NULL/-1) give the mesh the patch belongs to: faces
 * flagged is_interrior_level_boundary_{x,p} then take their fluxes from the finer patch (defined in the AMR part below). */
struct vo_mesh_s;
static void replace_level_boundary_fluxes_hook(void* M, const vo_fields* F, int p, int step);
static void fct_substep_impl(vo_patch* P, const vo_fields* F, double timestep, int step, int subStep, void* M, int pidx) {
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
        if (M) replace_level_boundary_fluxes_hook(M, F, pidx, step);
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

void vo_fct_substep(vo_patch* P, const vo_fields* F, double timestep, int step, int subStep) {
    fct_substep_impl(P, F, timestep, step, subStep, NULL, -1);
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
            if (!(nested && (nested[NS(P, i, j)] & 1))) {
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

/* =====================================================================================================
 * AMR: multi-patch, multi-level meshes.  One vo_mesh = one Mesh (one species).
 * Levels are indexed by depth, 0 = finest (Mesh.cpp:814).  Patch order inside a level is the caller's order =
 * the reference's Level::rectangles order; the connectivity tables depend on it (later calls overwrite).
 * ===================================================================================================== */
enum { VO_NESTED = 1, VO_LBX = 2, VO_LBP = 4 };
typedef struct {
    int depth, ns_x, ns_p;        /* strips per x side (n_p/r + 2) and per p side (n_x/r)  (Rectangle.cpp:40-48) */
    int* nb[4];                   /* 0 xm, 1 xp, 2 pm, 3 pp: neighbour patch index, -1 = the BoundaryCondition object */
    unsigned char* same[4];       /* is_exterrior_boundary_same_level_* */
    int *finer, *finer_x, *finer_p;   /* per padded cell: patch index or -1 (finer_level, finer_level_x, finer_level_p) */
    unsigned char* flags;         /* VO_NESTED | VO_LBX | VO_LBP per padded cell */
} vo_conn;
typedef struct {
    int n, r, n_levels;
    vo_patch** P;
    vo_conn* C;
    int* level_start;             /* n_levels + 1 offsets into order[] */
    int* order;                   /* patch indices grouped by depth, caller order inside a level */
    double coefs_ref[3 * 8];      /* interpolatCoefsREF (Rectangle.cpp:94-101) */
} vo_mesh;

static int on_line(int x, int y1, int y2) { return ((x - y1) > -1) && ((y2 - x) > -1); }   /* Rectangle.cpp:667-669 */
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* Rectangle::CalculateConnectivityFromFiner (Rectangle.cpp:768-864): ic = current (coarse), jf = rectangle (finer) */
static void conn_from_finer(vo_mesh* M, int ic, int jf) {
    const int r = M->r;
    vo_patch *Cc = M->P[ic], *Ff = M->P[jf];
    vo_conn *cc = &M->C[ic], *cf = &M->C[jf];
    const int n_x = Cc->n_x, n_p = Cc->n_p;
    int x_pos_2 = Ff->x_pos / r - Cc->x_pos, p_pos_2 = Ff->p_pos / r - Cc->p_pos;
    int n_x_2 = Ff->n_x / r, n_p_2 = Ff->n_p / r;
    int lower_x = imax(0, x_pos_2), upper_x = imin(n_x, x_pos_2 + n_x_2);
    int lower_p = imax(0, p_pos_2), upper_p = imin(n_p, p_pos_2 + n_p_2);
    for (int i = lower_x; i < upper_x; i++)
        for (int j = lower_p; j < upper_p; j++) { cc->flags[NS(Cc, i, j)] |= VO_NESTED; cc->finer[NS(Cc, i, j)] = jf; }
    if (((x_pos_2 + n_x_2) < (n_x + 1)) && ((x_pos_2 + n_x_2) > -1))
        for (int i = lower_p; i < upper_p; i++) { cc->finer_x[NS(Cc, x_pos_2 + n_x_2, i)] = jf; cc->flags[NS(Cc, x_pos_2 + n_x_2, i)] |= VO_LBX; }
    if (((x_pos_2 + n_x_2) < n_x) && ((x_pos_2 + n_x_2) > -1)) {
        for (int i = lower_p; i < upper_p; i++) { int j = i - p_pos_2 + 1; cf->nb[1][j] = ic; cf->same[1][j] = 0; }
        if (on_line(p_pos_2 - 1, 0, n_p - 1)) { cf->nb[1][0] = ic; cf->same[1][0] = 0; }
        if (on_line(p_pos_2 + n_p_2, 0, n_p - 1)) { cf->nb[1][cf->ns_x - 1] = ic; cf->same[1][cf->ns_x - 1] = 0; }
    }
    if ((x_pos_2 < (n_x + 1)) && (x_pos_2 > -1))
        for (int i = lower_p; i < upper_p; i++) { cc->finer_x[NS(Cc, x_pos_2, i)] = jf; cc->flags[NS(Cc, x_pos_2, i)] |= VO_LBX; }
    if ((x_pos_2 < (n_x + 1)) && (x_pos_2 > 0)) {
        for (int i = lower_p; i < upper_p; i++) { int j = i - p_pos_2 + 1; cf->nb[0][j] = ic; cf->same[0][j] = 0; }
        if (on_line(p_pos_2 - 1, 0, n_p - 1)) { cf->nb[0][0] = ic; cf->same[0][0] = 0; }
        if (on_line(p_pos_2 + n_p_2, 0, n_p - 1)) { cf->nb[0][cf->ns_x - 1] = ic; cf->same[0][cf->ns_x - 1] = 0; }
    }
    if (((p_pos_2 + n_p_2) < (n_p + 1)) && ((p_pos_2 + n_p_2) > -1))
        for (int i = lower_x; i < upper_x; i++) { cc->finer_p[NS(Cc, i, p_pos_2 + n_p_2)] = jf; cc->flags[NS(Cc, i, p_pos_2 + n_p_2)] |= VO_LBP; }
    if (((p_pos_2 + n_p_2) < n_p) && ((p_pos_2 + n_p_2) > -1))
        for (int i = lower_x; i < upper_x; i++) { int j = i - x_pos_2; cf->nb[3][j] = ic; cf->same[3][j] = 0; }
    if ((p_pos_2 < (n_p + 1)) && (p_pos_2 > -1))
        for (int i = lower_x; i < upper_x; i++) { cc->finer_p[NS(Cc, i, p_pos_2)] = jf; cc->flags[NS(Cc, i, p_pos_2)] |= VO_LBP; }
    if ((p_pos_2 < (n_p + 1)) && (p_pos_2 > 0))
        for (int i = lower_x; i < upper_x; i++) { int j = i - x_pos_2; cf->nb[2][j] = ic; cf->same[2][j] = 0; }
}

/* Rectangle::CalculateConnectivitySame (Rectangle.cpp:671-766): a = this, b = rectangle */
static void conn_same(vo_mesh* M, int a, int b) {
    const int r = M->r;
    vo_patch *A = M->P[a], *B = M->P[b];
    vo_conn* ca = &M->C[a];
    const int n_x = A->n_x, n_p = A->n_p, x_pos = A->x_pos, p_pos = A->p_pos;
    const int n_x_2 = B->n_x, n_p_2 = B->n_p, x_pos_2 = B->x_pos, p_pos_2 = B->p_pos;
    const int x_pos_r = x_pos_2 - x_pos, p_pos_r = p_pos_2 - p_pos;
    const int last = ca->ns_x - 1;
    if (x_pos_r == n_x) {
        int lower = imax(p_pos_2, p_pos), upper = imin(p_pos + n_p, p_pos_2 + n_p_2);
        if (on_line(-1, p_pos_r, p_pos_r + n_p_2 - 1)) { ca->nb[1][0] = b; ca->same[1][0] = 1; }
        if (on_line(n_p, p_pos_r, p_pos_r + n_p_2 - 1)) { ca->nb[1][last] = b; ca->same[1][last] = 1; }
        while (lower < upper) { int i1 = (lower - p_pos) / r + 1; ca->nb[1][i1] = b; ca->same[1][i1] = 1; lower += r; }
    }
    if (x_pos_r == -n_x_2) {
        int lower = imax(p_pos_2, p_pos), upper = imin(p_pos + n_p, p_pos_2 + n_p_2);
        if (on_line(-1, p_pos_r, p_pos_r + n_p_2 - 1)) { ca->nb[0][0] = b; ca->same[0][0] = 1; }
        if (on_line(n_p, p_pos_r, p_pos_r + n_p_2 - 1)) { ca->nb[0][last] = b; ca->same[0][last] = 1; }
        while (lower < upper) { int i1 = (lower - p_pos) / r + 1; ca->nb[0][i1] = b; ca->same[0][i1] = 1; lower += r; }
    }
    if (p_pos_r == n_p) {
        int lower = imax(x_pos_2, x_pos), upper = imin(x_pos + n_x, x_pos_2 + n_x_2);
        if (on_line(-1, x_pos_r, x_pos_r + n_x_2 - 1)) { ca->nb[0][last] = b; ca->same[0][last] = 1; }
        if (on_line(n_x, x_pos_r, x_pos_r + n_x_2 - 1)) { ca->nb[1][last] = b; ca->same[1][last] = 1; }
        while (lower < upper) { int i1 = (lower - x_pos) / r; ca->nb[3][i1] = b; ca->same[3][i1] = 1; lower += r; }
    }
    if (p_pos_r == -n_p_2) {
        int lower = imax(x_pos_2, x_pos), upper = imin(x_pos + n_x, x_pos_2 + n_x_2);
        if (on_line(-1, x_pos_r, x_pos_r + n_x_2 - 1)) { ca->nb[0][0] = b; ca->same[0][0] = 1; }
        if (on_line(n_x, x_pos_r, x_pos_r + n_x_2 - 1)) { ca->nb[1][0] = b; ca->same[1][0] = 1; }
        while (lower < upper) { int i1 = (lower - x_pos) / r; ca->nb[2][i1] = b; ca->same[2][i1] = 1; lower += r; }
    }
}

/* Rectangle ctor tables + the connectivity passes of Mesh::promoteHierarchyToMesh (Mesh.cpp:840-861) */
vo_mesh* vo_mesh_create(int n, int r, int n_levels, vo_patch** patches, const int* depth) {
    vo_mesh* M = (vo_mesh*)calloc(1, sizeof(vo_mesh));
    M->n = n; M->r = r; M->n_levels = n_levels;
    M->P = (vo_patch**)malloc(sizeof(vo_patch*) * n);
    M->C = (vo_conn*)calloc(n, sizeof(vo_conn));
    M->order = (int*)malloc(sizeof(int) * n);
    M->level_start = (int*)calloc(n_levels + 1, sizeof(int));
    interp_coefs(r, M->coefs_ref);
    int k = 0;
    for (int l = 0; l < n_levels; l++) {
        M->level_start[l] = k;
        for (int p = 0; p < n; p++) if (depth[p] == l) M->order[k++] = p;
    }
    M->level_start[n_levels] = k;
    for (int p = 0; p < n; p++) {
        vo_patch* P = patches[p]; vo_conn* c = &M->C[p];
        M->P[p] = P; c->depth = depth[p];
        c->ns_x = P->n_p / r + 2; c->ns_p = P->n_x / r;
        long npad = (long)(P->n_x + 4) * (P->n_p + 4);
        for (int s = 0; s < 4; s++) {
            int ns = s < 2 ? c->ns_x : c->ns_p;
            c->nb[s] = (int*)malloc(sizeof(int) * ns); c->same[s] = (unsigned char*)malloc(ns);
            for (int i = 0; i < ns; i++) { c->nb[s][i] = -1; c->same[s][i] = 1; }
        }
        c->finer = (int*)malloc(sizeof(int) * npad); c->finer_x = (int*)malloc(sizeof(int) * npad); c->finer_p = (int*)malloc(sizeof(int) * npad);
        for (long i = 0; i < npad; i++) { c->finer[i] = -1; c->finer_x[i] = -1; c->finer_p[i] = -1; }
        c->flags = (unsigned char*)calloc(npad, 1);
    }
    for (int l = 1; l < n_levels; l++)
        for (int a = M->level_start[l]; a < M->level_start[l + 1]; a++)
            for (int b = M->level_start[l - 1]; b < M->level_start[l]; b++) conn_from_finer(M, M->order[a], M->order[b]);
    for (int l = 0; l < n_levels; l++)
        for (int a = M->level_start[l]; a < M->level_start[l + 1]; a++)
            for (int b = M->level_start[l]; b < M->level_start[l + 1]; b++) if (a != b) conn_same(M, M->order[a], M->order[b]);
    return M;
}
void vo_mesh_destroy(vo_mesh* M) {
    for (int p = 0; p < M->n; p++) {
        vo_conn* c = &M->C[p];
        for (int s = 0; s < 4; s++) { free(c->nb[s]); free(c->same[s]); }
        free(c->finer); free(c->finer_x); free(c->finer_p); free(c->flags);
    }
    free(M->P); free(M->C); free(M->order); free(M->level_start); free(M);
}
/* test access to the derived tables: side 0..3 -> nb/same strips; returns the strip count */
int vo_mesh_get_strips(const vo_mesh* M, int p, int side, int* nb, unsigned char* same) {
    const vo_conn* c = &M->C[p]; int ns = side < 2 ? c->ns_x : c->ns_p;
    for (int i = 0; i < ns; i++) { nb[i] = c->nb[side][i]; same[i] = c->same[side][i]; }
    return ns;
}
void vo_mesh_get_flags(const vo_mesh* M, int p, unsigned char* flags) {
    const vo_patch* P = M->P[p]; memcpy(flags, M->C[p].flags, (size_t)(P->n_x + 4) * (P->n_p + 4));
}

static inline double* fstate(vo_patch* P, int val) { return val == 2 ? P->f2 : (val == 1 ? P->f1 : P->f0); }
/* Rectangle::GetValueFromSameLevel (Rectangle.cpp:307-312); nb < 0 = BoundaryCondition (BoundaryCondition.cpp:6-8) */
static double same_level_value(vo_mesh* M, int nb, int i, int j, int val) {
    if (nb < 0) return 0.0;
    vo_patch* Q = M->P[nb];
    return fstate(Q, val)[NS(Q, i - Q->x_pos, j - Q->p_pos)];
}
/* Rectangle::GetValueFromFinerLevel (Rectangle.cpp:314-327) on the finer patch nb */
static double finer_level_value(vo_mesh* M, int nb, int i, int j, int val) {
    vo_patch* Q = M->P[nb]; const int r = M->r;
    int i_f = i * r - Q->x_pos, j_f = j * r - Q->p_pos;
    double t = 0.0;
    for (int k = 0; k < r; k++) for (int l = 0; l < r; l++) t += fstate(Q, val)[NS(Q, i_f + k, j_f + l)];
    t /= pow((double)r, 2.0);
    return t;
}
/* Rectangle::GetWenoValueFromCoarseLevel (Rectangle.cpp:343-415) on the coarse patch nb; out has r*r (d = -1) or 2r values */
static void coarse_level_values(vo_mesh* M, int nb, int i, int j, int d, int val, double* out) {
    vo_patch* Q = M->P[nb]; const int r = M->r;
    const double* f = fstate(Q, val);
    int ic = i / r - Q->x_pos, jc = j / r - Q->p_pos;
    double temps[5][8], ip[64], part[8], sum = 0.0;
    for (int k = -2; k < 3; k++)
        interpolants(M->coefs_ref, r, f[NS(Q, ic - 2, jc + k)], f[NS(Q, ic - 1, jc + k)], f[NS(Q, ic, jc + k)], f[NS(Q, ic + 1, jc + k)], f[NS(Q, ic + 2, jc + k)], temps[k + 2]);
    for (int k = 0; k < r; k++) {
        interpolants(M->coefs_ref, r, temps[0][k], temps[1][k], temps[2][k], temps[3][k], temps[4][k], part);
        for (int l = 0; l < r; l++) { ip[k * r + l] = part[l]; sum += part[l]; }
    }
    double correction = f[NS(Q, ic, jc)] - 1.0 / pow((double)r, 2) * sum;
    for (int k = 0; k < r * r; k++) ip[k] += correction;
    if (d == 0) for (int k = 0; k < r; k++) { out[2 * k] = ip[r * k + r - 1]; out[2 * k + 1] = ip[r * k + r - 2]; }
    else if (d == 1) for (int k = 0; k < r; k++) { out[2 * k] = ip[k]; out[2 * k + 1] = ip[r + k]; }
    else if (d == 2) for (int k = 0; k < r; k++) { out[2 * k] = ip[r * k]; out[2 * k + 1] = ip[r * k + 1]; }
    else if (d == 3) for (int k = 0; k < r; k++) { out[2 * k] = ip[r * (r - 1) + k]; out[2 * k + 1] = ip[r * (r - 2) + k]; }
    else for (int k = 0; k < r * r; k++) out[k] = ip[k];
}
/* Rectangle::UpdateInterriorPoints (Rectangle.cpp:329-337) */
static void update_interior_points(vo_mesh* M, int p, int val) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p]; double* f = fstate(P, val);
    for (int i = 0; i < P->n_x; i++)
        for (int j = 0; j < P->n_p; j++)
            if (c->flags[NS(P, i, j)] & VO_NESTED) f[NS(P, i, j)] = finer_level_value(M, c->finer[NS(P, i, j)], P->x_pos + i, P->p_pos + j, val);
}
/* Rectangle::UpdateSameLevelBoundaries (Rectangle.cpp:562-614) */
static void update_same_level_boundaries(vo_mesh* M, int p, int val) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p]; double* f = fstate(P, val); const int r = M->r;
    const int n_x = P->n_x, n_p = P->n_p, x_pos = P->x_pos, p_pos = P->p_pos;
    for (int i = 0; i < c->ns_x - 2; i++) if (c->same[0][i + 1]) for (int j = 0; j < r; j++) {
        f[NS(P, -1, i * r + j)] = same_level_value(M, c->nb[0][i + 1], x_pos - 1, p_pos + i * r + j, val);
        f[NS(P, -2, i * r + j)] = same_level_value(M, c->nb[0][i + 1], x_pos - 2, p_pos + i * r + j, val);
    }
    for (int i = 0; i < c->ns_x - 2; i++) if (c->same[1][i + 1]) for (int j = 0; j < r; j++) {
        f[NS(P, n_x, i * r + j)] = same_level_value(M, c->nb[1][i + 1], x_pos + n_x, p_pos + i * r + j, val);
        f[NS(P, n_x + 1, i * r + j)] = same_level_value(M, c->nb[1][i + 1], x_pos + n_x + 1, p_pos + i * r + j, val);
    }
    for (int i = 0; i < c->ns_p; i++) if (c->same[2][i]) for (int j = 0; j < r; j++) {
        f[NS(P, i * r + j, -1)] = same_level_value(M, c->nb[2][i], x_pos + i * r + j, p_pos - 1, val);
        f[NS(P, i * r + j, -2)] = same_level_value(M, c->nb[2][i], x_pos + i * r + j, p_pos - 2, val);
    }
    for (int i = 0; i < c->ns_p; i++) if (c->same[3][i]) for (int j = 0; j < r; j++) {
        f[NS(P, i * r + j, n_p)] = same_level_value(M, c->nb[3][i], x_pos + i * r + j, p_pos + n_p, val);
        f[NS(P, i * r + j, n_p + 1)] = same_level_value(M, c->nb[3][i], x_pos + i * r + j, p_pos + n_p + 1, val);
    }
}
/* Rectangle::UpdateCornerPoints (Rectangle.cpp:1130-1214) = the corner blocks of UpdateDifferentLevelBoundaries (484-560) */
static void update_corner_points(vo_mesh* M, int p, int val) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p]; double* f = fstate(P, val); const int r = M->r;
    const int n_x = P->n_x, x_pos = P->x_pos, p_pos = P->p_pos;
    double temp[16];
    for (int side = 0; side < 2; side++) {            /* 0: xm (ghost columns -1,-2), 1: xp (n_x, n_x+1) */
        const int g1 = side == 0 ? -1 : n_x, g2 = side == 0 ? -2 : n_x + 1;
        const int q1 = side == 0 ? x_pos - 1 : x_pos + n_x, q2 = side == 0 ? x_pos - 2 : x_pos + n_x + 1;
        const int d = side == 0 ? 3 : 1;
        for (int which = 0; which < 2; which++) {     /* 0: lower corner (strip -1), 1: upper corner */
            const int i = which == 0 ? -1 : c->ns_x - 2;
            const int j0 = which == 0 ? r - 2 : 0, j1 = which == 0 ? r : 2;
            const int e = i + 1;
            if (!c->same[side][e]) {
                coarse_level_values(M, c->nb[side][e], q1, p_pos + i * r, d, val, temp);
                for (int j = j0; j < j1; j++) { f[NS(P, g1, i * r + j)] = temp[2 * j]; f[NS(P, g2, i * r + j)] = temp[2 * j + 1]; }
            } else {
                for (int j = j0; j < j1; j++) {
                    f[NS(P, g1, i * r + j)] = same_level_value(M, c->nb[side][e], q1, p_pos + i * r + j, val);
                    f[NS(P, g2, i * r + j)] = same_level_value(M, c->nb[side][e], q2, p_pos + i * r + j, val);
                }
            }
        }
    }
}
/* Rectangle::UpdateDifferentLevelBoundaries (Rectangle.cpp:417-560) */
static void update_different_level_boundaries(vo_mesh* M, int p, int val) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p]; double* f = fstate(P, val); const int r = M->r;
    const int n_x = P->n_x, n_p = P->n_p, x_pos = P->x_pos, p_pos = P->p_pos;
    double temp[16];
    for (int i = 0; i < c->ns_x - 2; i++) if (!c->same[0][i + 1]) {
        coarse_level_values(M, c->nb[0][i + 1], x_pos - 1, p_pos + i * r, 3, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, -1, i * r + j)] = temp[2 * j]; f[NS(P, -2, i * r + j)] = temp[2 * j + 1]; }
    }
    for (int i = 0; i < c->ns_x - 2; i++) if (!c->same[1][i + 1]) {
        coarse_level_values(M, c->nb[1][i + 1], x_pos + n_x, p_pos + i * r, 1, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, n_x, i * r + j)] = temp[2 * j]; f[NS(P, n_x + 1, i * r + j)] = temp[2 * j + 1]; }
    }
    for (int i = 0; i < c->ns_p; i++) if (!c->same[2][i]) {
        coarse_level_values(M, c->nb[2][i], x_pos + i * r, p_pos - 1, 0, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, i * r + j, -1)] = temp[2 * j]; f[NS(P, i * r + j, -2)] = temp[2 * j + 1]; }
    }
    for (int i = 0; i < c->ns_p; i++) if (!c->same[3][i]) {
        coarse_level_values(M, c->nb[3][i], x_pos + i * r, p_pos + n_p, 2, val, temp);
        for (int j = 0; j < r; j++) { f[NS(P, i * r + j, n_p)] = temp[2 * j]; f[NS(P, i * r + j, n_p + 1)] = temp[2 * j + 1]; }
    }
    update_corner_points(M, p, val);
}
/* Level::PushData (Level.cpp:88-126), update types 0..3 */
static void level_push(vo_mesh* M, int l, int type, int val) {
    for (int a = M->level_start[l]; a < M->level_start[l + 1]; a++) {
        int p = M->order[a];
        if (type == 0) update_interior_points(M, p, val);
        else if (type == 1) update_same_level_boundaries(M, p, val);
        else if (type == 2) update_different_level_boundaries(M, p, val);
        else update_corner_points(M, p, val);
    }
}
/* Mesh::PushData (Mesh.cpp:91-106) */
void vo_mesh_push_data(vo_mesh* M, int val) {
    const int nl = M->n_levels;
    level_push(M, 0, 1, val);
    for (int i = 1; i < nl; i++) { level_push(M, i, 0, val); level_push(M, i, 1, val); }
    level_push(M, nl - 1, 3, val);
    for (int i = nl - 1; i > 0; i--) level_push(M, i - 1, 2, val);
}

/* ---- coarse-fine flux matching: RGKGetFlux{X,P,XL,PL} + CalculateFluxToCoarse* (Rectangle.cpp:1032-1098, 1216-1253;
 * Rectangle.hpp:132-178).  Recursive through the levels, recomputed from the finer patch's f1. */
static double rgk_flux(vo_mesh* M, const vo_fields* F, int p, int i, int j, int kind);   /* kind 0 X, 1 P, 2 XL, 3 PL */
static double flux_to_coarse(vo_mesh* M, const vo_fields* F, int p, int i, int j, int kind) {
    vo_patch* P = M->P[p]; const int r = M->r;
    i = i * r - P->x_pos; j = j * r - P->p_pos;
    double t = 0.0;
    for (int k = 0; k < r; k++) t += (kind == 0 || kind == 2) ? rgk_flux(M, F, p, i, j + k, kind) : rgk_flux(M, F, p, i + k, j, kind);
    t *= (1.0 / (r * r));
    return t;
}
static double rgk_flux(vo_mesh* M, const vo_fields* F, int p, int i, int j, int kind) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p];
    const long idx = NS(P, i, j);
    const int isx = (kind == 0 || kind == 2);
    if (c->flags[idx] & (isx ? VO_LBX : VO_LBP))
        return flux_to_coarse(M, F, isx ? c->finer_x[idx] : c->finer_p[idx], P->x_pos + i, P->p_pos + j, kind);
    const double w3 = 1.0 / 48.0, dx_inv = 1 / P->dx, dp_inv = 1 / P->dp, q = P->q, cc = CS * CS * P->m;
    const double* f = P->f1;
    if (kind == 0) {
        double as = q * q * a_sq(F, finest_index(P, i));
        double am = dp_inv * cc * (gamma_(P, momentum(P, j + 1), as) - gamma_(P, momentum(P, j), as));
        double ap1 = dp_inv * cc * (gamma_(P, momentum(P, j + 2.0), as) - gamma_(P, momentum(P, j + 1.0), as));
        double am1 = dp_inv * cc * (gamma_(P, momentum(P, j), as) - gamma_(P, momentum(P, j - 1.0), as));
        double fm = vo_weno(f[NS(P, i - 2, j)], f[NS(P, i - 1, j)], f[NS(P, i, j)], f[NS(P, i + 1, j)], am > 0.0);
        double fp1 = vo_weno(f[NS(P, i - 2, j + 1)], f[NS(P, i - 1, j + 1)], f[NS(P, i, j + 1)], f[NS(P, i + 1, j + 1)], ap1 > 0.0);
        double fm1 = vo_weno(f[NS(P, i - 2, j - 1)], f[NS(P, i - 1, j - 1)], f[NS(P, i, j - 1)], f[NS(P, i + 1, j - 1)], am1 > 0.0);
        return dx_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
    } else if (kind == 2) {
        double as = q * q * a_sq(F, finest_index(P, i));
        double am = dp_inv * cc * (gamma_(P, momentum(P, j + 1), as) - gamma_(P, momentum(P, j), as));
        return dx_inv * (am > 0.0 ? f[NS(P, i - 1, j)] : f[NS(P, i, j)]) * am;
    }
    double as_1 = q * q * a_sq(F, finest_index(P, i)), as_2 = q * q * a_sq(F, finest_index(P, i + 1));
    double Em = q * patch_efield(P, F, i), mom = momentum(P, j);
    double am = Em - cc * dx_inv * (gamma_(P, mom, as_2) - gamma_(P, mom, as_1));
    if (kind == 3) return dp_inv * (am > 0.0 ? f[NS(P, i, j - 1)] : f[NS(P, i, j)]) * am;
    double as_0 = q * q * a_sq(F, finest_index(P, i - 1)), as_3 = q * q * a_sq(F, finest_index(P, i + 2));
    double Ep1 = q * patch_efield(P, F, i + 1), Em1 = q * patch_efield(P, F, i - 1);
    double ap1 = Ep1 - cc * dx_inv * (gamma_(P, mom, as_3) - gamma_(P, mom, as_2));
    double am1 = Em1 - cc * dx_inv * (gamma_(P, mom, as_1) - gamma_(P, mom, as_0));
    double fm = vo_weno(f[NS(P, i, j - 2)], f[NS(P, i, j - 1)], f[NS(P, i, j)], f[NS(P, i, j + 1)], am > 0.0);
    double fp1 = vo_weno(f[NS(P, i + 1, j - 2)], f[NS(P, i + 1, j - 1)], f[NS(P, i + 1, j)], f[NS(P, i + 1, j + 1)], ap1 > 0.0);
    double fm1 = vo_weno(f[NS(P, i - 1, j - 2)], f[NS(P, i - 1, j - 1)], f[NS(P, i - 1, j)], f[NS(P, i - 1, j + 1)], am1 > 0.0);
    return dp_inv * (fm * am + w3 * (fp1 - fm1) * (ap1 - am1));
}
/* the four flux loops of sub-step 0 at faces flagged is_interrior_level_boundary_{x,p} (Rectangle.cpp:1313-1394):
 * called between the unflagged flux evaluation and the RK combination */
static void replace_level_boundary_fluxes(vo_mesh* M, const vo_fields* F, int p, int step) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p];
    const int nx = P->n_x, np = P->n_p; const long npad = (long)(nx + 4) * (np + 4);
    double* FxHs = P->FxH + step * npad; double* FpHs = P->FpH + step * npad;
    for (int i = 0; i < nx + 1; i++)
        for (int j = -1; j < np + 1; j++) {
            long idx = NS(P, i, j);
            if (c->flags[idx] & VO_LBX) {
                FxHs[idx] = flux_to_coarse(M, F, c->finer_x[idx], P->x_pos + i, P->p_pos + j, 0);
                if (step == 0) P->FxL[idx] = flux_to_coarse(M, F, c->finer_x[idx], P->x_pos + i, P->p_pos + j, 2);
            }
        }
    for (int i = -1; i < nx + 1; i++)
        for (int j = 0; j < np + 1; j++) {
            long idx = NS(P, i, j);
            if (c->flags[idx] & VO_LBP) {
                FpHs[idx] = flux_to_coarse(M, F, c->finer_p[idx], P->x_pos + i, P->p_pos + j, 1);
                if (step == 0) P->FpL[idx] = flux_to_coarse(M, F, c->finer_p[idx], P->x_pos + i, P->p_pos + j, 3);
            }
        }
}
static void replace_level_boundary_fluxes_hook(void* M, const vo_fields* F, int p, int step) {
    replace_level_boundary_fluxes((vo_mesh*)M, F, p, step);
}
/* Level::FCTTimeStep (Level.cpp:12-17) */
void vo_mesh_substep(vo_mesh* M, const vo_fields* F, int depth, double dt, int step, int subStep) {
    for (int a = M->level_start[depth]; a < M->level_start[depth + 1]; a++)
        fct_substep_impl(M->P[M->order[a]], F, dt, step, subStep, M, M->order[a]);
}

/* ---- limiter sync: Mesh::PushBoundaryC (Mesh.cpp:904-917) ------------------------------------------------- */
/* Rectangle::SetCFromSameLevel (Rectangle.cpp:1795-1833): n = this (the neighbour), c = rectangle (the caller) */
static void set_c_from_same_level(vo_mesh* M, int n, int i, int j, int c, int t) {
    if (n < 0) return;   /* BoundaryCondition: value-initialised, refinementRatio = 0 -> zero-trip loops (quirk Q8) */
    vo_patch *N = M->P[n], *Cl = M->P[c]; const int r = M->r;
    int in = i - N->x_pos, jn = j - N->p_pos, icl = i - Cl->x_pos, jcl = j - Cl->p_pos;
    for (int k = 0; k < r; k++) {
        if (t == 0) {
            double cx = N->FxDS[NS(N, in, jn + k)] > 0.0 ? fmin(Cl->Rp[NS(Cl, icl, jcl + k)], N->Rm[NS(N, in - 1, jn + k)])
                                                          : fmin(N->Rp[NS(N, in - 1, jn + k)], Cl->Rm[NS(Cl, icl, jcl + k)]);
            N->Cx[NS(N, in, jn + k)] = cx; Cl->Cx[NS(Cl, icl, jcl + k)] = cx;
        } else if (t == 1) {
            double cx = N->FxDS[NS(N, in, jn + k)] > 0.0 ? fmin(Cl->Rm[NS(Cl, icl - 1, jcl + k)], N->Rp[NS(N, in, jn + k)])
                                                          : fmin(N->Rm[NS(N, in, jn + k)], Cl->Rp[NS(Cl, icl - 1, jcl + k)]);
            N->Cx[NS(N, in, jn + k)] = cx; Cl->Cx[NS(Cl, icl, jcl + k)] = cx;
        } else if (t == 2) {
            double cp = N->FpDS[NS(N, in + k, jn)] > 0.0 ? fmin(Cl->Rp[NS(Cl, icl + k, jcl)], N->Rm[NS(N, in + k, jn - 1)])
                                                          : fmin(N->Rp[NS(N, in + k, jn - 1)], Cl->Rm[NS(Cl, icl + k, jcl)]);
            N->Cp[NS(N, in + k, jn)] = cp; Cl->Cp[NS(Cl, icl + k, jcl)] = cp;
        } else {
            double cp = N->FpDS[NS(N, in + k, jn)] > 0.0 ? fmin(Cl->Rm[NS(Cl, icl + k, jcl - 1)], N->Rp[NS(N, in + k, jn)])
                                                          : fmin(N->Rm[NS(N, in + k, jn)], Cl->Rp[NS(Cl, icl + k, jcl - 1)]);
            N->Cp[NS(N, in + k, jn)] = cp; Cl->Cp[NS(Cl, icl + k, jcl)] = cp;
        }
    }
}
/* Rectangle::UpdateCFromSameLevel (Rectangle.cpp:1879-1917) */
static void update_c_from_same_level(vo_mesh* M, int n, int i, int j, int c, int t) {
    if (n < 0) return;
    vo_patch *N = M->P[n], *Cl = M->P[c]; const int r = M->r;
    int in = i - N->x_pos, jn = j - N->p_pos, icl = i - Cl->x_pos, jcl = j - Cl->p_pos;
    for (int k = 0; k < r; k++) {
        if (t < 2) {
            double cx = fmin(N->Cx[NS(N, in, jn + k)], Cl->Cx[NS(Cl, icl, jcl + k)]);
            N->Cx[NS(N, in, jn + k)] = cx; Cl->Cx[NS(Cl, icl, jcl + k)] = cx;
        } else {
            double cp = fmin(N->Cp[NS(N, in + k, jn)], Cl->Cp[NS(Cl, icl + k, jcl)]);
            N->Cp[NS(N, in + k, jn)] = cp; Cl->Cp[NS(Cl, icl + k, jcl)] = cp;
        }
    }
}
/* Rectangle::SetCFromDifferentLevel (Rectangle.cpp:1625-1704): n = this (coarse), c = rectangle (the fine caller) */
static void set_c_from_different_level(vo_mesh* M, int n, int i, int j, int c, int t) {
    vo_patch *N = M->P[n], *Cl = M->P[c]; const int r = M->r;
    int icl = i - Cl->x_pos, jcl = j - Cl->p_pos, ico = i / r - N->x_pos, jco = j / r - N->p_pos;
    if (t == 0) {
        double cx = N->Cx[NS(N, ico, jco)];
        for (int k = 0; k < r; k++) cx = fmin(cx, Cl->FxDS[NS(Cl, icl, jcl + k)] > 0.0 ? Cl->Rp[NS(Cl, icl, jcl + k)] : Cl->Rm[NS(Cl, icl, jcl + k)]);
        cx = fmin(cx, N->FxDS[NS(N, ico, jco)] > 0.0 ? N->Rm[NS(N, ico - 1, jco)] : N->Rp[NS(N, ico - 1, jco)]);
        for (int k = 0; k < r; k++) Cl->Cx[NS(Cl, icl, jcl + k)] = cx;
        N->Cx[NS(N, ico, jco)] = cx;
    } else if (t == 1) {
        double cx = N->Cx[NS(N, ico, jco)];
        for (int k = 0; k < r; k++) cx = fmin(cx, Cl->FxDS[NS(Cl, icl, jcl + k)] > 0.0 ? Cl->Rm[NS(Cl, icl - 1, jcl + k)] : Cl->Rp[NS(Cl, icl - 1, jcl + k)]);
        cx = fmin(cx, N->FxDS[NS(N, ico, jco)] > 0.0 ? N->Rp[NS(N, ico, jco)] : N->Rm[NS(N, ico, jco)]);
        for (int k = 0; k < r; k++) Cl->Cx[NS(Cl, icl, jcl + k)] = cx;
        N->Cx[NS(N, ico, jco)] = cx;
    } else if (t == 2) {
        double cp = N->Cp[NS(N, ico, jco)];
        for (int k = 0; k < r; k++) cp = fmin(cp, Cl->FpDS[NS(Cl, icl + k, jcl)] > 0.0 ? Cl->Rp[NS(Cl, icl + k, jcl)] : Cl->Rm[NS(Cl, icl + k, jcl)]);
        cp = fmin(cp, N->FpDS[NS(N, ico, jco)] > 0.0 ? N->Rm[NS(N, ico, jco - 1)] : N->Rp[NS(N, ico, jco - 1)]);
        for (int k = 0; k < r; k++) Cl->Cp[NS(Cl, icl + k, jcl)] = cp;
        N->Cp[NS(N, ico, jco)] = cp;
    } else {
        double cp = N->Cp[NS(N, ico, jco)];
        for (int k = 0; k < r; k++) cp = fmin(cp, Cl->FpDS[NS(Cl, icl + k, jcl)] > 0.0 ? Cl->Rm[NS(Cl, icl + k, jcl - 1)] : Cl->Rp[NS(Cl, icl + k, jcl - 1)]);
        cp = fmin(cp, N->FpDS[NS(N, ico, jco)] > 0.0 ? N->Rp[NS(N, ico, jco)] : N->Rm[NS(N, ico, jco)]);
        for (int k = 0; k < r; k++) Cl->Cp[NS(Cl, icl + k, jcl)] = cp;
        N->Cp[NS(N, ico, jco)] = cp;
    }
}
/* CalculateSameBoundaryC / CalculateDifferentBoundaryC / UpdateSameBoundaryC (Rectangle.cpp:1835-1877, 1706-1763, 1919-1960) */
static void boundary_c_pass(vo_mesh* M, int p, int pass) {
    vo_patch* P = M->P[p]; vo_conn* c = &M->C[p]; const int r = M->r;
    const int n_x = P->n_x, n_p = P->n_p, x_pos = P->x_pos, p_pos = P->p_pos;
    for (int side = 0; side < 4; side++) {
        const int ns = side < 2 ? c->ns_x - 2 : c->ns_p;
        for (int i = 0; i < ns; i++) {
            const int e = side < 2 ? i + 1 : i;
            const int same = c->same[side][e], nb = c->nb[side][e];
            int gi, gj;
            if (side == 0) { gi = x_pos; gj = p_pos + r * i; }
            else if (side == 1) { gi = x_pos + n_x; gj = p_pos + r * i; }
            else if (side == 2) { gi = x_pos + r * i; gj = p_pos; }
            else { gi = x_pos + r * i; gj = p_pos + n_p; }
            if (pass == 4 && same) set_c_from_same_level(M, nb, gi, gj, p, side);
            else if (pass == 5 && !same) set_c_from_different_level(M, nb, gi, gj, p, side);
            else if (pass == 6 && same) update_c_from_same_level(M, nb, gi, gj, p, side);
        }
    }
}
void vo_mesh_push_boundary_c(vo_mesh* M) {
    for (int pass = 4; pass <= 6; pass++)
        for (int l = 0; l < M->n_levels; l++)
            for (int a = M->level_start[l]; a < M->level_start[l + 1]; a++) boundary_c_pass(M, M->order[a], pass);
}
/* Mesh::Advance (Mesh.cpp:64-89) */
void vo_mesh_advance(vo_mesh* M, const vo_fields* F, double dt, int step) {
    const int nl = M->n_levels;
    for (int i = nl - 1; i > -1; i--) vo_mesh_substep(M, F, i, dt, step, 0);
    vo_mesh_push_data(M, 2);
    for (int i = nl - 1; i > -1; i--) vo_mesh_substep(M, F, i, dt, step, 1);
    vo_mesh_push_boundary_c(M);
    for (int i = nl - 1; i > -1; i--) vo_mesh_substep(M, F, i, dt, step, 2);
    vo_mesh_push_data(M, 1);
    if (step == 5) for (int i = nl - 1; i > -1; i--) vo_mesh_substep(M, F, i, dt, step, 3);
}
/* Mesh::InterpolateRhoAndJToFinestMesh (Mesh.cpp:52-56) -> Level::CollectRhoAndJ (Level.cpp:19-62): adds this species'
 * moments to charge_s[N] and J[N].  scratch: 4*N doubles. */
void vo_mesh_moments(vo_mesh* M, const vo_fields* F, double* charge_s, double* J, double* scratch) {
    const int N = F->x_size;
    double *chargeL = scratch, *currentL = scratch + N, *cR = scratch + 2 * N, *jR = scratch + 3 * N;
    for (int l = 0; l < M->n_levels; l++) {
        for (int i = 0; i < N; i++) { chargeL[i] = 0.0; currentL[i] = 0.0; }
        for (int a = M->level_start[l]; a < M->level_start[l + 1]; a++) {
            int p = M->order[a]; vo_patch* P = M->P[p];
            vo_patch_moments(P, F, M->C[p].flags, cR, jR);
            int n = P->n_x * P->rtb, shift = P->x_pos * P->rtb;
            for (int j = 0; j < n; j++) { chargeL[shift + j] += cR[j]; currentL[shift + j] += jR[j]; }
        }
        for (int i = 0; i < N; i++) { charge_s[i] += chargeL[i]; J[i] += currentL[i]; }
    }
}
