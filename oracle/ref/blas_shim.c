/*
 * blas_shim.c -- naive column-major dgemm_/dgemv_/daxpy_ for the oracle build
 * of the FemTech reference.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference links an external, unversioned BLAS (CMakeLists.txt:140).  The
 * hot path only calls dgemm on 3x3 (and 24x24x3 for the mass matrix) and dgemv
 * on 6x24, i.e. dot products of <= 6 terms.  This shim fixes the summation
 * order to plain left-to-right, which is what oracle/femtech_oracle.c restates;
 * "reference result" in this repository means "reference sources + this shim".
 */
#include <stdio.h>
#include <stdlib.h>

void dgemm_(char *ta, char *tb, int *m, int *n, int *k, double *alpha,
            double *a, int *lda, double *b, int *ldb, double *beta, double *c,
            int *ldc) {
  int TA = (*ta == 'T' || *ta == 't'), TB = (*tb == 'T' || *tb == 't');
  for (int j = 0; j < *n; j++)
    for (int i = 0; i < *m; i++) {
      double s = 0.0;
      for (int l = 0; l < *k; l++) {
        double av = TA ? a[l + i * (*lda)] : a[i + l * (*lda)];
        double bv = TB ? b[j + l * (*ldb)] : b[l + j * (*ldb)];
        s += av * bv;
      }
      if (*beta == 0.0)
        c[i + j * (*ldc)] = (*alpha) * s;
      else
        c[i + j * (*ldc)] = (*beta) * c[i + j * (*ldc)] + (*alpha) * s;
    }
}

void dgemv_(char *t, int *m, int *n, double *alpha, double *a, int *lda,
            double *x, int *incx, double *beta, double *y, int *incy) {
  int T = (*t == 'T' || *t == 't');
  int ly = T ? *n : *m, lx = T ? *m : *n;
  for (int i = 0; i < ly; i++) {
    double s = 0.0;
    for (int j = 0; j < lx; j++) {
      double av = T ? a[j + i * (*lda)] : a[i + j * (*lda)];
      s += av * x[j * (*incx)];
    }
    if (*beta == 0.0)
      y[i * (*incy)] = (*alpha) * s;
    else
      y[i * (*incy)] = (*beta) * y[i * (*incy)] + (*alpha) * s;
  }
}

void daxpy_(int *n, double *al, double *x, int *ix, double *y, int *iy) {
  for (int i = 0; i < *n; i++) y[i * (*iy)] += (*al) * x[i * (*ix)];
}

/* LAPACK is only reached from the implicit toy solvers (out of scope). */
void dgesv_(void) { fprintf(stderr, "dgesv_: not in oracle shim\n"); abort(); }
void dgetrf_(void) { fprintf(stderr, "dgetrf_: not in oracle shim\n"); abort(); }
void dgetrs_(void) { fprintf(stderr, "dgetrs_: not in oracle shim\n"); abort(); }
