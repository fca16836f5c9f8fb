// Test infrastructure (oracle build only): minimal stand-in for Intel MKL's <mkl.h> so the
// UNMODIFIED reference sources under /root/reference compile with -DUSINGMKL.
// LAPACKE_dgetrf/dgetrs forward to the Fortran LAPACK in OpenBLAS 0.3.15; vdLn/vdExp are libm loops.
#pragma once
#include <cmath>
typedef int lapack_int;
#define LAPACK_COL_MAJOR 102
extern "C" void dgetrf_(const int*, const int*, double*, const int*, int*, int*);
extern "C" void dgetrs_(const char*, const int*, const int*, const double*, const int*, const int*, double*, const int*, int*);
static inline lapack_int LAPACKE_dgetrf(int, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv) {
    int info = 0; dgetrf_(&m, &n, a, &lda, ipiv, &info); return info; }
static inline lapack_int LAPACKE_dgetrs(int, char trans, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                                        const lapack_int* ipiv, double* b, lapack_int ldb) {
    int info = 0; dgetrs_(&trans, &n, &nrhs, a, &lda, ipiv, b, &ldb, &info); return info; }
static inline void vdLn(int n, const double* a, double* y) { for (int i = 0; i < n; i++) y[i] = std::log(a[i]); }
static inline void vdExp(int n, const double* a, double* y) { for (int i = 0; i < n; i++) y[i] = std::exp(a[i]); }
