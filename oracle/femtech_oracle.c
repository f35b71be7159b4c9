/*
 * femtech_oracle.c -- TEST INFRASTRUCTURE ONLY (see femtech_oracle.h).
 *
 * Plain-C restatement of the FemTech hex8 explicit-dynamics hot path.  Every
 * function follows the reference's arithmetic operation by operation (same
 * summation order, same divisions vs reciprocal multiplies, same quirks) so
 * that, compiled with -ffp-contract=off, it reproduces a reference build made
 * with the same flag BIT FOR BIT (tests/test_oracle_golden.py).
 *
 * Citations are relative to /root/reference.
 */
#include "femtech_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NDIM 3
static const double huge_dt = 1e20; /* include/GlobalVariables.h:16 */

/* Mixed C3D8 / C3D4 meshes (ShapeFunctions.cpp:71-164): gpoff[e] = Gauss points before element e (8 per hexahedron,
 * 1 per tetrahedron); NULL = all hexahedra.  With h hexahedra and t tetrahedra before e: gpoff = 8h + t, e = h + t. */
static inline int el_hex_before(const oracle_state *s, int e) { return s->gpoff ? (s->gpoff[e] - e) / 7 : e; }
static inline int NGP(const oracle_state *s, int e) { return s->gpoff ? s->gpoff[e + 1] - s->gpoff[e] : 8; }
static inline int NEN(const oracle_state *s, int e) { return NGP(s, e) == 8 ? 8 : 4; }
static inline int GP0(const oracle_state *s, int e) { return s->gpoff ? s->gpoff[e] : 8 * e; }          /* gpPtr, detFptr */
static inline int C0(const oracle_state *s, int e) { return 4 * e + 4 * el_hex_before(s, e); }           /* eptr */
static inline int SHP0(const oracle_state *s, int e) { return 4 * e + 60 * el_hex_before(s, e); }        /* gptr; dsptr = 3x */

/* ------------------------------------------------------------------------ */
/* tiny BLAS restatements: naive left-to-right sums, identical to            */
/* oracle/ref/blas_shim.c (the definition of "reference result")  */
/* ------------------------------------------------------------------------ */
/* C(3x3) = alpha * op(A) * op(B), column-major, beta = 0 */
static void mm3(int ta, int tb, double alpha, const double *a, const double *b,
                double *c) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
      for (int l = 0; l < 3; ++l) {
        double av = ta ? a[l + 3 * i] : a[i + 3 * l];
        double bv = tb ? b[j + 3 * l] : b[l + 3 * j];
        s += av * bv;
      }
      c[i + 3 * j] = alpha * s;
    }
}

/* src/math/math.cpp:8-28 */
static void inverse3x3Matrix(const double *mat, double *invMat, double *det) {
  double detLocal = mat[0] * (mat[4] * mat[8] - mat[5] * mat[7]) -
                    mat[3] * (mat[1] * mat[8] - mat[7] * mat[2]) +
                    mat[6] * (mat[1] * mat[5] - mat[4] * mat[2]);
  double invdet = 1 / detLocal;
  invMat[0] = (mat[4] * mat[8] - mat[5] * mat[7]) * invdet;
  invMat[3] = (mat[6] * mat[5] - mat[3] * mat[8]) * invdet;
  invMat[6] = (mat[3] * mat[7] - mat[6] * mat[4]) * invdet;
  invMat[1] = (mat[7] * mat[2] - mat[1] * mat[8]) * invdet;
  invMat[4] = (mat[0] * mat[8] - mat[6] * mat[2]) * invdet;
  invMat[7] = (mat[1] * mat[6] - mat[0] * mat[7]) * invdet;
  invMat[2] = (mat[1] * mat[5] - mat[2] * mat[4]) * invdet;
  invMat[5] = (mat[2] * mat[3] - mat[0] * mat[5]) * invdet;
  invMat[8] = (mat[0] * mat[4] - mat[1] * mat[3]) * invdet;
  *det = detLocal;
}

/* src/math/math.cpp:30-41 (ndim == 3 branch) */
static double normOfCrossProduct(const double *a, const double *b) {
  double z = a[0] * b[1] - a[1] * b[0];
  double x = a[1] * b[2] - a[2] * b[1];
  double y = -a[0] * b[2] + a[2] * b[0];
  return sqrt(x * x + y * y + z * z);
}

/* src/math/math.cpp:44-48 */
static double tripleProduct(const double *s, const double *a, const double *b) {
  return s[2] * (a[0] * b[1] - a[1] * b[0]) +
         s[0] * (a[1] * b[2] - a[2] * b[1]) -
         s[1] * (a[0] * b[2] - a[2] * b[0]);
}

/* ------------------------------------------------------------------------ */
/* Geometry (stable time step)                                               */
/* ------------------------------------------------------------------------ */
/* src/math/Geometry.cpp:3-27 */
double oracle_volumeHexahedron(const double *L) {
  double q0[3], q1[3], q2[3], q3[3], q4[3], q5[3];
  for (int i = 0; i < 3; ++i) {
    q0[i] = L[0 + i] - L[3 + i] + L[6 + i] - L[9 + i] + L[12 + i] - L[15 + i] +
            L[18 + i] - L[21 + i];
    q1[i] = L[0 + i] - L[3 + i] - L[6 + i] + L[9 + i] - L[12 + i] + L[15 + i] +
            L[18 + i] - L[21 + i];
    q2[i] = -L[0 + i] + L[3 + i] + L[6 + i] - L[9 + i] - L[12 + i] + L[15 + i] +
            L[18 + i] - L[21 + i];
    q3[i] = L[0 + i] + L[3 + i] - L[6 + i] - L[9 + i] - L[12 + i] - L[15 + i] +
            L[18 + i] + L[21 + i];
    q4[i] = -L[0 + i] - L[3 + i] + L[6 + i] + L[9 + i] - L[12 + i] - L[15 + i] +
            L[18 + i] + L[21 + i];
    q5[i] = -L[0 + i] - L[3 + i] - L[6 + i] - L[9 + i] + L[12 + i] + L[15 + i] +
            L[18 + i] + L[21 + i];
  }
  return (tripleProduct(q0, q4, q3) + tripleProduct(q2, q0, q1) +
          tripleProduct(q1, q3, q5)) / 192.0 +
         tripleProduct(q2, q4, q5) / 64.0;
}

/* src/math/Geometry.cpp:66-76 */
static double volumeTetrahedron(const double *L) {
  double a[3], b[3], c[3];
  for (int i = 0; i < 3; ++i) {
    a[i] = L[i] - L[9 + i];
    b[i] = L[3 + i] - L[9 + i];
    c[i] = L[6 + i] - L[9 + i];
  }
  return fabs(tripleProduct(a, b, c)) / 6.0;
}

/* src/math/Geometry.cpp:29-64, including the signed (no fabs) parallelogram
 * test of :46 */
double oracle_areaHexahedronFace(const double *L, const int *index) {
  const double *p1 = &L[3 * index[0]];
  const double *p2 = &L[3 * index[1]];
  const double *p3 = &L[3 * index[2]];
  const double *p4 = &L[3 * index[3]];
  double tol = 1e-6;
  double centerD[3], c1[3], c2[3];
  for (int i = 0; i < 3; ++i) {
    centerD[i] = 0.25 * (p1[i] - p2[i] + p3[i] - p4[i]);
    c1[i] = 0.25 * (-p1[i] + p2[i] + p3[i] - p4[i]);
    c2[i] = 0.25 * (-p1[i] - p2[i] + p3[i] + p4[i]);
  }
  if ((centerD[0] < tol) && (centerD[1] < tol) && (centerD[2] < tol)) {
    return 4.0 * normOfCrossProduct(c1, c2);
  }
  double t = sqrt(3.0) / 3.0;
  const double q[2] = {-t, t};
  double area = 0.0;
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 2; ++j) {
      double v1[3], v2[3];
      for (int k = 0; k < 3; ++k) {
        v1[k] = q[j] * centerD[k] + c1[k];
        v2[k] = q[i] * centerD[k] + c2[k];
      }
      area += normOfCrossProduct(v1, v2);
    }
  }
  return area;
}

/* src/elements/CharacteristicLength/CalculateCharacteristicLength_C3D8.cpp:3-30 */
static double characteristicLength_C3D8(const oracle_state *s, int e) {
  double ec[24];
  for (int j = 0; j < 8; ++j) {
    int n = s->connectivity[C0(s, e) + j];
    for (int k = 0; k < 3; ++k) {
      int index = NDIM * n + k;
      ec[j * NDIM + k] = s->coordinates[index] + s->displacements[index];
    }
  }
  static const int index[24] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 3, 7, 4,
                                1, 2, 6, 5, 0, 1, 5, 4, 3, 2, 6, 7};
  double cl = oracle_volumeHexahedron(ec);
  double areaMax = 0.0, faceArea;
  for (int i = 0; i < 6; ++i) {
    faceArea = oracle_areaHexahedronFace(ec, &index[i * 4]);
    if (faceArea > areaMax) areaMax = faceArea;
  }
  return cl / areaMax;
}

/* src/elements/CharacteristicLength/CalculateCharacteristicLength_C3D4.cpp:5-47: smallest altitude */
static double distancePointPlane(const double *x0, const double *x1, const double *x2, const double *x3) {
  double v1[3], v2[3], normal[3];
  for (int i = 0; i < 3; ++i) {
    v1[i] = x2[i] - x1[i];
    v2[i] = x3[i] - x1[i];
  }
  normal[0] = v1[1] * v2[2] - v1[2] * v2[1];
  normal[1] = -v1[0] * v2[2] + v1[2] * v2[0];
  normal[2] = v1[0] * v2[1] - v1[1] * v2[0];
  double crossNorm = sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
  for (int i = 0; i < 3; ++i) normal[i] /= crossNorm;
  double v3[3];
  for (int i = 0; i < 3; ++i) v3[i] = x0[i] - x1[i];
  return fabs(normal[0] * v3[0] + normal[1] * v3[1] + normal[2] * v3[2]);
}
static double characteristicLength_C3D4(const oracle_state *s, int e) {
  static const int index[16] = {0, 1, 2, 3, 1, 2, 3, 0, 2, 3, 0, 1, 3, 0, 1, 2};
  double ec[12];
  for (int j = 0; j < 4; ++j) {
    int n = s->connectivity[C0(s, e) + j];
    for (int k = 0; k < 3; ++k) ec[j * NDIM + k] = s->coordinates[NDIM * n + k] + s->displacements[NDIM * n + k];
  }
  double altitudeMin = 1e6, minAlt;
  for (int i = 0; i < 4; ++i) {
    minAlt = distancePointPlane(&ec[index[4 * i] * NDIM], &ec[index[4 * i + 1] * NDIM], &ec[index[4 * i + 2] * NDIM],
                                &ec[index[4 * i + 3] * NDIM]);
    if (minAlt < altitudeMin) altitudeMin = minAlt;
  }
  return altitudeMin;
}

/* src/timestep/CalculateTimeStep.cpp:7-21 */
double oracle_CalculateTimeStep(const oracle_state *s, int e) {
  double le = NEN(s, e) == 8 ? characteristicLength_C3D8(s, e) : characteristicLength_C3D4(s, e);
  int pide = s->pid[e];
  double mu = s->properties[ORACLE_MAXMATPARAMS * pide + 1];
  double lambda = s->properties[ORACLE_MAXMATPARAMS * pide + 2];
  double rho = s->properties[ORACLE_MAXMATPARAMS * pide + 0];
  double nu = 0.5 * lambda / (lambda + mu);
  double ce = sqrt(lambda * (1.0 / nu - 1.0) / rho);
  return le / ce;
}

/* src/timestep/StableTimeStep.cpp:11-30 */
double oracle_StableTimeStep_local(const oracle_state *s) {
  double dtMin = huge_dt;
  for (int i = 0; i < s->nElements; i++) {
    int isNotRigid = 0;
    for (int j = C0(s, i); j < C0(s, i) + NEN(s, i); ++j) {
      int index = s->connectivity[j] * NDIM;
      if (!(s->boundary[index] & s->boundary[index + 1] &
            s->boundary[index + 2])) {
        isNotRigid = 1;
        break;
      }
    }
    if (isNotRigid) {
      double dtElem = oracle_CalculateTimeStep(s, i);
      if (dtElem < dtMin) dtMin = dtElem;
    }
  }
  return dtMin;
}

/* ------------------------------------------------------------------------ */
/* Shape functions (reference configuration, once)                           */
/* ------------------------------------------------------------------------ */
/* src/fem/ShapeFunctions/GaussQuadrature3D.cpp:17-58 */
static const double GP_A = 0.577350269189626;
static const double gp_sign[8][3] = {{-1, -1, 1}, {1, -1, 1}, {1, 1, 1},
                                     {-1, 1, 1},  {-1, -1, -1}, {1, -1, -1},
                                     {1, 1, -1},  {-1, 1, -1}};

/* src/fem/ShapeFunctions/ShapeFunction_C3D8.cpp:4-128 */
static void shapeFunction_C3D8(oracle_state *s, int e, int gp) {
  const double chi = gp_sign[gp][0] * GP_A;
  const double eta = gp_sign[gp][1] * GP_A;
  const double iota = gp_sign[gp][2] * GP_A;
  double *shp = &s->shp[SHP0(s, e) + 8 * gp];
  double *d = &s->dshp[3 * SHP0(s, e) + 24 * gp];
  const double *X = s->coordinates;
  const int *c = &s->connectivity[C0(s, e)];

  shp[0] = ((1 - chi) * (1 - eta) * (1 - iota)) / 8;
  shp[1] = ((1 + chi) * (1 - eta) * (1 - iota)) / 8;
  shp[2] = ((1 + chi) * (1 + eta) * (1 - iota)) / 8;
  shp[3] = ((1 - chi) * (1 + eta) * (1 - iota)) / 8;
  shp[4] = ((1 - chi) * (1 - eta) * (1 + iota)) / 8;
  shp[5] = ((1 + chi) * (1 - eta) * (1 + iota)) / 8;
  shp[6] = ((1 + chi) * (1 + eta) * (1 + iota)) / 8;
  shp[7] = ((1 - chi) * (1 + eta) * (1 + iota)) / 8;

  d[3 * 0 + 0] = -((eta - 1) * (iota - 1)) / 8;
  d[3 * 1 + 0] = ((eta - 1) * (iota - 1)) / 8;
  d[3 * 2 + 0] = -((eta + 1) * (iota - 1)) / 8;
  d[3 * 3 + 0] = ((eta + 1) * (iota - 1)) / 8;
  d[3 * 4 + 0] = ((eta - 1) * (iota + 1)) / 8;
  d[3 * 5 + 0] = -((eta - 1) * (iota + 1)) / 8;
  d[3 * 6 + 0] = ((eta + 1) * (iota + 1)) / 8;
  d[3 * 7 + 0] = -((eta + 1) * (iota + 1)) / 8;

  d[3 * 0 + 1] = -((chi - 1) * (iota - 1)) / 8;
  d[3 * 1 + 1] = ((chi + 1) * (iota - 1)) / 8;
  d[3 * 2 + 1] = -((chi + 1) * (iota - 1)) / 8;
  d[3 * 3 + 1] = ((chi - 1) * (iota - 1)) / 8;
  d[3 * 4 + 1] = ((chi - 1) * (iota + 1)) / 8;
  d[3 * 5 + 1] = -((chi + 1) * (iota + 1)) / 8;
  d[3 * 6 + 1] = ((chi + 1) * (iota + 1)) / 8;
  d[3 * 7 + 1] = -((chi - 1) * (iota + 1)) / 8;

  d[3 * 0 + 2] = -((chi - 1) * (eta - 1)) / 8;
  d[3 * 1 + 2] = ((chi + 1) * (eta - 1)) / 8;
  d[3 * 2 + 2] = -((chi + 1) * (eta + 1)) / 8;
  d[3 * 3 + 2] = ((chi - 1) * (eta + 1)) / 8;
  d[3 * 4 + 2] = ((chi - 1) * (eta - 1)) / 8;
  d[3 * 5 + 2] = -((chi + 1) * (eta - 1)) / 8;
  d[3 * 6 + 2] = ((chi + 1) * (eta + 1)) / 8;
  d[3 * 7 + 2] = -((chi - 1) * (eta + 1)) / 8;

  /* Jacobian from four paired differences per entry (:78-93) */
  double xs[9];
#define XC(n, j) X[NDIM * c[n] + (j)]
  for (int j = 0; j < NDIM; j++) {
    xs[0 + j * NDIM] = (XC(1, j) - XC(0, j)) * d[3 * 1 + 0] +
                       (XC(2, j) - XC(3, j)) * d[3 * 2 + 0] +
                       (XC(5, j) - XC(4, j)) * d[3 * 5 + 0] +
                       (XC(6, j) - XC(7, j)) * d[3 * 6 + 0];
    xs[1 + j * NDIM] = (XC(2, j) - XC(1, j)) * d[3 * 2 + 1] +
                       (XC(3, j) - XC(0, j)) * d[3 * 3 + 1] +
                       (XC(6, j) - XC(5, j)) * d[3 * 6 + 1] +
                       (XC(7, j) - XC(4, j)) * d[3 * 7 + 1];
    xs[2 + j * NDIM] = (XC(4, j) - XC(0, j)) * d[3 * 4 + 2] +
                       (XC(5, j) - XC(1, j)) * d[3 * 5 + 2] +
                       (XC(6, j) - XC(2, j)) * d[3 * 6 + 2] +
                       (XC(7, j) - XC(3, j)) * d[3 * 7 + 2];
  }
#undef XC
  double det, J_Inv[9];
  inverse3x3Matrix(xs, J_Inv, &det);
  s->detJacobian[GP0(s, e) + gp] = det;
  /* dN/dX (:108-119) */
  for (int i = 0; i < 8; ++i) {
    double *b = &d[3 * i];
    double c1 = b[0] * J_Inv[0] + b[1] * J_Inv[3] + b[2] * J_Inv[6];
    double c2 = b[0] * J_Inv[1] + b[1] * J_Inv[4] + b[2] * J_Inv[7];
    double c3 = b[0] * J_Inv[2] + b[1] * J_Inv[5] + b[2] * J_Inv[8];
    b[0] = c1;
    b[1] = c2;
    b[2] = c3;
  }
}

/* src/fem/ShapeFunctions/ShapeFunctions.cpp:181-252 */
/* src/fem/ShapeFunctions/ShapeFunction_C3D4.cpp:6-85: one Gauss point at (1/4,1/4,1/4), weight 1/6
 * (GaussQuadrature3D.cpp:62-69), detJ = |det| (:70) */
static void shapeFunction_C3D4(oracle_state *s, int e) {
  const double chi = 0.25, eta = 0.25, iota = 0.25;
  double *shp = &s->shp[SHP0(s, e)];
  double *d = &s->dshp[3 * SHP0(s, e)];
  const double *X = s->coordinates;
  const int *c = &s->connectivity[C0(s, e)];
  shp[0] = chi;
  shp[1] = eta;
  shp[2] = iota;
  shp[3] = 1.0 - eta - iota - chi;
  d[3 * 0 + 0] = 1.0; d[3 * 1 + 0] = 0.0; d[3 * 2 + 0] = 0.0; d[3 * 3 + 0] = -1.0;
  d[3 * 0 + 1] = 0.0; d[3 * 1 + 1] = 1.0; d[3 * 2 + 1] = 0.0; d[3 * 3 + 1] = -1.0;
  d[3 * 0 + 2] = 0.0; d[3 * 1 + 2] = 0.0; d[3 * 2 + 2] = 1.0; d[3 * 3 + 2] = -1.0;
  double xs[9];
  for (int j = 0; j < 3; ++j) {
    xs[0 + 3 * j] = X[NDIM * c[0] + j] - X[NDIM * c[3] + j];
    xs[1 + 3 * j] = X[NDIM * c[1] + j] - X[NDIM * c[3] + j];
    xs[2 + 3 * j] = X[NDIM * c[2] + j] - X[NDIM * c[3] + j];
  }
  double det, J_Inv[9];
  inverse3x3Matrix(xs, J_Inv, &det);
  s->detJacobian[GP0(s, e)] = fabs(det);
  s->gaussWeights[GP0(s, e)] = 1.0 / 6.0;
  for (int i = 0; i < 4; ++i) {
    double *b = &d[3 * i];
    double c1 = b[0] * J_Inv[0] + b[1] * J_Inv[3] + b[2] * J_Inv[6];
    double c2 = b[0] * J_Inv[1] + b[1] * J_Inv[4] + b[2] * J_Inv[7];
    double c3 = b[0] * J_Inv[2] + b[1] * J_Inv[5] + b[2] * J_Inv[8];
    b[0] = c1;
    b[1] = c2;
    b[2] = c3;
  }
}

/* src/fem/ShapeFunctions/ShapeFunctions.cpp:181-252 */
void oracle_ShapeFunctions(oracle_state *s) {
  const int nE = s->nElements;
  const size_t nGP = (size_t)GP0(s, nE), nShp = (size_t)SHP0(s, nE);
  memset(s->shp, 0, sizeof(double) * nShp);
  memset(s->dshp, 0, sizeof(double) * 3 * nShp);
  memset(s->F, 0, sizeof(double) * 9 * nGP);
  memset(s->pk2, 0, sizeof(double) * 6 * nGP);
  for (size_t i = 0; i < nGP; ++i) {
    s->F[i * 9] = 1.0;
    s->F[i * 9 + 4] = 1.0;
    s->F[i * 9 + 8] = 1.0;
    s->detF[i] = 1.0;
    s->gaussWeights[i] = 1.0;
  }
  for (int e = 0; e < nE; ++e) {
    if (NEN(s, e) == 8) {
      for (int k = 0; k < 8; ++k) shapeFunction_C3D8(s, e, k);
    } else {
      shapeFunction_C3D4(s, e);
    }
  }
  if (s->Hn_1) memset(s->Hn_1, 0, sizeof(double) * 9 * nGP);
  if (s->Hn_2) memset(s->Hn_2, 0, sizeof(double) * 9 * nGP);
  if (s->S0n) memset(s->S0n, 0, sizeof(double) * 9 * nGP);
}

/* ------------------------------------------------------------------------ */
/* Lumped mass                                                               */
/* ------------------------------------------------------------------------ */
/* src/fem/Mass/Mass3D.cpp:5-67 + :127-157.  The 24x24 consistent matrix
 * N^T N has the block form Me[(3n+a) + 24*(3m+b)] = delta_ab * N_n * N_m; the
 * reference accumulates it with dgemm (k = 3 inner terms, two of which are
 * exact zeros), scales by w*detJ per Gauss point, by rho at the end, and then
 * lumps rows in ascending column-block order.  The exact zeros do not change
 * any rounding, so only the non-zero entries are carried here. */
void oracle_AssembleLumpedMass_local(oracle_state *s) {
  for (int e = 0; e < s->nElements; ++e) {
    double Me[8][8]; /* Me[n][m] = sum_gp (N_n N_m) * pre */
    memset(Me, 0, sizeof(Me));
    const int nen = NEN(s, e), ngp = NGP(s, e);
    for (int k = 0; k < ngp; ++k) {
      const double *shp = &s->shp[SHP0(s, e) + nen * k];
      int wIndex = GP0(s, e) + k;
      const double preFactor = s->gaussWeights[wIndex] * s->detJacobian[wIndex];
      for (int n = 0; n < nen; ++n)
        for (int m = 0; m < nen; ++m) {
          /* dgemm inner sum: 0 + Nn*Nm + (exact zeros) ; alpha = 1 */
          double MeGQ = shp[n] * shp[m];
          Me[n][m] += MeGQ * preFactor;
        }
    }
    double rho = s->properties[ORACLE_MAXMATPARAMS * s->pid[e]];
    for (int n = 0; n < nen; ++n)
      for (int m = 0; m < nen; ++m) Me[n][m] *= rho;
    /* row-sum lumping: Me[j] += Me[j + i*24], i = 1..23 (:140-144).  For row
     * j = 3n+a only columns i = 3m+a are non-zero; adding exact zeros is a
     * no-op, so the sum runs over m = 0..7 in ascending order starting from
     * the m = 0 column (i = a is the first non-zero column for row 3n+a:
     * for a > 0 the row starts from the zero in column 0). */
    for (int l = 0; l < nen; ++l) {
      const int gIndex = s->connectivity[C0(s, e) + l];
      for (int a = 0; a < 3; ++a) {
        double lumped = (a == 0) ? Me[l][0] : 0.0;
        for (int m = (a == 0) ? 1 : 0; m < nen; ++m) lumped += Me[l][m];
        s->mass[gIndex * NDIM + a] += lumped;
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Materials                                                                 */
/* ------------------------------------------------------------------------ */
/* src/math/InverseF.cpp:38-66 (ndim == 3) */
static void InverseF(const double *Fm, double detA, double *fInv) {
  double A[3][3];
  A[0][0] = Fm[0]; A[0][1] = Fm[1]; A[0][2] = Fm[2];
  A[1][0] = Fm[3]; A[1][1] = Fm[4]; A[1][2] = Fm[5];
  A[2][0] = Fm[6]; A[2][1] = Fm[7]; A[2][2] = Fm[8];
  fInv[0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / detA;
  fInv[1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / detA;
  fInv[2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / detA;
  fInv[3] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / detA;
  fInv[4] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / detA;
  fInv[5] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / detA;
  fInv[6] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / detA;
  fInv[7] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / detA;
  fInv[8] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / detA;
}

static void storeVoigt(const double *S, double *pk2) {
  pk2[0] = S[0]; pk2[1] = S[4]; pk2[2] = S[8];
  pk2[3] = S[7]; pk2[4] = S[6]; pk2[5] = S[3];
}

/* src/materials/CompressibleNeoHookean.cpp:14-58 */
static void CompressibleNeoHookean(const double *Fg, double J, const double *p,
                                   double *pk2) {
  double mu = p[1], lambda = p[2];
  double Cmat[9], Cinv[9], Cdet;
  mm3(1, 0, 1.0, Fg, Fg, Cmat);
  inverse3x3Matrix(Cmat, Cinv, &Cdet);
  double logJ = lambda * log(J);
  pk2[0] = mu * (1.0 - Cinv[0]) + logJ * Cinv[0];
  pk2[1] = mu * (1.0 - Cinv[4]) + logJ * Cinv[4];
  pk2[2] = mu * (1.0 - Cinv[8]) + logJ * Cinv[8];
  pk2[3] = -mu * Cinv[7] + logJ * Cinv[7];
  pk2[4] = -mu * Cinv[6] + logJ * Cinv[6];
  pk2[5] = -mu * Cinv[3] + logJ * Cinv[3];
}

/* src/materials/StVenantKirchhoff.cpp:25-48 */
static void StVenantKirchhoff(const double *Fg, const double *p, double *pk2) {
  double mu = p[1], lambda = p[2];
  double E[9], S[9];
  double half = 0.5;
  mm3(1, 0, half, Fg, Fg, E);
  E[0] -= half; E[4] -= half; E[8] -= half;
  double traceE = E[0] + E[4] + E[8];
  for (int i = 0; i < 9; ++i) S[i] = 2.0 * mu * E[i];
  S[0] += lambda * traceE;
  S[4] += lambda * traceE;
  S[8] += lambda * traceE;
  storeVoigt(S, pk2);
}

/* src/materials/LinearElastic.cpp:30-63 */
static void LinearElastic(const double *Fg, double J, const double *p,
                          double *pk2) {
  double mu = p[1], lambda = p[2];
  double eps[9], P[9], fInv[9], S[9];
  const double trEps = Fg[0] + Fg[4] + Fg[8] - 3.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int indexL = j + i * 3;
      eps[indexL] = (Fg[indexL] + Fg[i + j * 3]);
    }
  eps[0] -= 2.0; eps[4] -= 2.0; eps[8] -= 2.0;
  for (int i = 0; i < 9; ++i) P[i] = mu * eps[i];
  P[0] += lambda * trEps;
  P[4] += lambda * trEps;
  P[8] += lambda * trEps;
  InverseF(Fg, J, fInv);
  mm3(0, 0, 1.0, fInv, P, S);
  storeVoigt(S, pk2);
}

/* Shared front half of HGOIsotropic.cpp:44-90 and
 * HGOIsotropicViscoelastic.cpp:54-100.  visco selects the rounding order of
 * the prefactor (A.4 of SURVEY.md): material 4 scales by
 * totalPrefactor = Jm23*(mu+fiber)/J, material 5 by Jm23*((mu+fiber)/J) one
 * multiply at a time. */
static void hgo_S(const double *Fg, double J, const double *p, int visco,
                  double *fInv, double *S) {
  const double mu = p[1], lambda = p[2], k1 = p[3], k2 = p[4];
  const double kappa = 1.0 / 3.0;
  const double K = lambda + 2.0 * mu / 3.0;
  const double hydroDiag = 0.5 * K * (J * J - 1.0) / J;
  double Bmat[9], STemp[9];
  mm3(0, 1, 1.0, Fg, Fg, Bmat);
  const double traceB = Bmat[0] + Bmat[4] + Bmat[8];
  const double Jm23 = pow(J, -2.0 / 3.0);
  const double I1 = Jm23 * traceB;
  const double E_alpha = kappa * (I1 - 3.0);
  double fiberPrefactor = 0.0;
  if (E_alpha > 0.0) {
    fiberPrefactor = 2.0 * k1 * exp(k2 * E_alpha * E_alpha) * E_alpha * kappa;
  }
  const double traceBby3 = traceB / 3.0;
  Bmat[0] = Bmat[0] - traceBby3;
  Bmat[4] = Bmat[4] - traceBby3;
  Bmat[8] = Bmat[8] - traceBby3;
  if (!visco) {
    const double totalPrefactor = Jm23 * (mu + fiberPrefactor) / J;
    for (int i = 0; i < 9; ++i) Bmat[i] = Bmat[i] * totalPrefactor;
  } else {
    const double totalPrefactor = (mu + fiberPrefactor) / J;
    for (int i = 0; i < 9; ++i) Bmat[i] = Bmat[i] * Jm23 * totalPrefactor;
  }
  Bmat[0] += hydroDiag;
  Bmat[4] += hydroDiag;
  Bmat[8] += hydroDiag;
  InverseF(Fg, J, fInv);
  mm3(0, 0, 1.0, fInv, Bmat, STemp);
  mm3(0, 1, J, STemp, fInv, S);
}

/* src/materials/HGOIsotropic.cpp:21-115 */
static void HGOIsotropic(const double *Fg, double J, const double *p,
                         double *pk2) {
  double fInv[9], S[9];
  hgo_S(Fg, J, p, 0, fInv, S);
  storeVoigt(S, pk2);
}

/* src/materials/HGOIsotropicViscoelastic.cpp:27-168 */
static void HGOIsotropicViscoelastic(const double *Fg, double J,
                                     const double *p, double dt, double *Hn_1,
                                     double *Hn_2, double *S0n, double *pk2) {
  const double g1 = p[5], t1 = p[6], g2 = p[7], t2 = p[8];
  double fInv[9], S[9], Cmat[9], Sic[9], Sdev[9];
  hgo_S(Fg, J, p, 1, fInv, S);
  mm3(1, 0, 1.0, Fg, Fg, Cmat);
  double SddC = 0.0;
  for (int i = 0; i < 9; ++i) SddC += Cmat[i] * S[i];
  SddC = SddC / 3.0;
  mm3(0, 1, SddC, fInv, fInv, Sic);
  for (int i = 0; i < 9; ++i) Sdev[i] = S[i] - Sic[i];
  const double rt1 = dt / t1;
  const double rt2 = dt / t2;
  const double c11 = exp(-rt1);
  const double c12 = exp(-rt2);
  const double c21 = g1 * (1 - c11) / rt1;
  const double c22 = g2 * (1 - c12) / rt2;
  for (int i = 0; i < 9; ++i) {
    Hn_1[i] = c11 * Hn_1[i] + c21 * (Sdev[i] - S0n[i]);
    Hn_2[i] = c12 * Hn_2[i] + c22 * (Sdev[i] - S0n[i]);
    S[i] = S[i] + Hn_1[i] + Hn_2[i];
  }
  storeVoigt(S, pk2);
  for (int i = 0; i < 9; ++i) S0n[i] = Sdev[i];
}

/* ------------------------------------------------------------------------ */
/* GetForce                                                                  */
/* ------------------------------------------------------------------------ */
/* src/elements/ElementCalculations/StrainDisplacementMatrix.cpp:25-77 */
static void StrainDisplacementMatrix(const double *dn, const double *Fg,
                                     double *Bmat) {
  double dnIdx = dn[0], dnIdy = dn[1], dnIdz = dn[2];
  double F11 = Fg[0], F21 = Fg[1], F31 = Fg[2];
  double F12 = Fg[3], F22 = Fg[4], F32 = Fg[5];
  double F13 = Fg[6], F23 = Fg[7], F33 = Fg[8];
  double B11 = dnIdx * F11, B12 = dnIdx * F21, B13 = dnIdx * F31;
  double B21 = dnIdy * F12, B22 = dnIdy * F22, B23 = dnIdy * F32;
  double B31 = dnIdz * F13, B32 = dnIdz * F23, B33 = dnIdz * F33;
  double B41 = dnIdy * F13 + dnIdz * F12;
  double B42 = dnIdy * F23 + dnIdz * F22;
  double B43 = dnIdy * F33 + dnIdz * F32;
  double B51 = dnIdx * F13 + dnIdz * F11;
  double B52 = dnIdx * F23 + dnIdz * F21;
  double B53 = dnIdx * F33 + dnIdz * F31;
  double B61 = dnIdx * F12 + dnIdy * F11;
  double B62 = dnIdx * F22 + dnIdy * F21;
  double B63 = dnIdx * F32 + dnIdy * F31;
  Bmat[0] = B11; Bmat[1] = B21; Bmat[2] = B31; Bmat[3] = B41; Bmat[4] = B51; Bmat[5] = B61;
  Bmat[6] = B12; Bmat[7] = B22; Bmat[8] = B32; Bmat[9] = B42; Bmat[10] = B52; Bmat[11] = B62;
  Bmat[12] = B13; Bmat[13] = B23; Bmat[14] = B33; Bmat[15] = B43; Bmat[16] = B53; Bmat[17] = B63;
}

/* src/fem/SolidMechanics/GetForce_3D.cpp:11-46 */
int oracle_GetForce_local(oracle_state *s) {
  const int nDOF = NDIM * s->nNodes;
  memcpy(s->f_net, s->fe, nDOF * sizeof(double));
  memset(s->fi, 0, nDOF * sizeof(double));
  int bad = 0;
  for (int e = 0; e < s->nElements; e++) {
    double fintLocal[24];
    memset(fintLocal, 0, sizeof(fintLocal));
    const int *conn = &s->connectivity[C0(s, e)];
    const int pide = s->pid[e];
    const double *props = &s->properties[ORACLE_MAXMATPARAMS * pide];
    const int nen = NEN(s, e), ngp = NGP(s, e), gp0 = GP0(s, e);
    for (int gp = 0; gp < ngp; gp++) {
      double *Fg = &s->F[9 * (gp0 + gp)];
      const double *dshp = &s->dshp[3 * SHP0(s, e) + 3 * nen * gp];
      /* CalculateDeformationGradient.cpp:11-25 */
      for (int i = 0; i < NDIM; i++)
        for (int j = 0; j < NDIM; j++) {
          double theSum = 0.0;
          for (int k = 0; k < nen; k++) {
            int node_a = conn[k];
            theSum = theSum + (s->coordinates[NDIM * node_a + i] +
                               s->displacements[NDIM * node_a + i]) *
                                  dshp[k * NDIM + j];
          }
          Fg[NDIM * j + i] = theSum;
        }
      /* DeterminateF.cpp:43-56 */
      {
        double da = Fg[0], db = Fg[3], dc = Fg[6];
        double dd = Fg[1], de = Fg[4], df = Fg[7];
        double dg = Fg[2], dh = Fg[5], di = Fg[8];
        s->detF[gp0 + gp] = da * (de * di - df * dh) -
                              db * (dd * di - df * dg) +
                              dc * (dd * dh - de * dg);
      }
      const double J = s->detF[gp0 + gp];
      double *pk2 = &s->pk2[6 * (gp0 + gp)];
      /* StressUpdate.cpp:7-27 */
      switch (s->materialID[pide]) {
        case 0: break;
        case 1: CompressibleNeoHookean(Fg, J, props, pk2); break;
        case 2: StVenantKirchhoff(Fg, props, pk2); break;
        case 3: LinearElastic(Fg, J, props, pk2); break;
        case 4: HGOIsotropic(Fg, J, props, pk2); break;
        case 5:
          HGOIsotropicViscoelastic(Fg, J, props, s->dt,
                                   &s->Hn_1[9 * (gp0 + gp)],
                                   &s->Hn_2[9 * (gp0 + gp)],
                                   &s->S0n[9 * (gp0 + gp)], pk2);
          break;
        default: bad = 1; break;
      }
      /* InternalForceUpdate.cpp:4-28: B (6x24), fintGQ = B^T sigma via the
       * naive dgemv ('T', m=6, n=24, alpha=1, beta=0), then the weighted add */
      double B[144];
      for (int k = 0; k < nen; ++k)
        StrainDisplacementMatrix(&dshp[3 * k], Fg, &B[18 * k]);
      const int wIndex = gp0 + gp;
      const double preFactor = s->gaussWeights[wIndex] * s->detJacobian[wIndex];
      for (int k = 0; k < 3 * nen; ++k) {
        double sum = 0.0;
        for (int j = 0; j < 6; ++j) sum += B[j + 6 * k] * pk2[j];
        double fintGQ = 1.0 * sum;
        fintLocal[k] += preFactor * fintGQ;
      }
    }
    /* scatter (:39-44) */
    for (int k = 0; k < nen; ++k) {
      int dIndex = conn[k];
      for (int l = 0; l < NDIM; ++l)
        s->fi[dIndex * NDIM + l] += fintLocal[k * NDIM + l];
    }
  }
  return bad;
}

/* src/fem/SolidMechanics/GetForce_3D.cpp:49-51 */
void oracle_GetForce_finish(oracle_state *s) {
  const int nDOF = NDIM * s->nNodes;
  for (int i = 0; i < nDOF; ++i) s->f_net[i] -= s->fi[i];
}

/* src/fem/SolidMechanics/CalculateAcclerations.cpp:7-11 */
void oracle_CalculateAccelerations(oracle_state *s) {
  const int nDOF = NDIM * s->nNodes;
  for (int i = 0; i < nDOF; ++i)
    if (!s->boundary[i]) s->accelerations[i] = s->f_net[i] / s->mass[i];
}

/* src/fem/SolidMechanics/GetForce_3D.cpp:54-102 and src/fem/Mass/Mass3D.cpp:
 * 77-125 for ranks emulated in one process.  Every rank packs first (the
 * reference posts all sends before any add), then every rank adds what its
 * neighbours packed, neighbour by neighbour in its own list order. */
void oracle_halo_sum(oracle_state **ranks, int nranks, int field) {
  double **sendbuf = (double **)calloc(nranks, sizeof(double *));
  for (int r = 0; r < nranks; ++r) {
    oracle_state *s = ranks[r];
    double *a = field == 0 ? s->fi : s->mass;
    int total = s->sendProcessCount ? s->sendNeighbourCountCum[s->sendProcessCount] : 0;
    sendbuf[r] = (double *)malloc(sizeof(double) * (NDIM * (size_t)total + 1));
    for (int i = 0; i < total; ++i)
      memcpy(&sendbuf[r][NDIM * i], &a[NDIM * s->sendNodeIndex[i]],
             sizeof(double) * NDIM);
  }
  for (int r = 0; r < nranks; ++r) {
    oracle_state *s = ranks[r];
    double *a = field == 0 ? s->fi : s->mass;
    for (int p = 0; p < s->sendProcessCount; ++p) {
      int q = s->sendProcessID[p];
      oracle_state *o = ranks[q];
      /* find this rank in the neighbour's list */
      int po = -1;
      for (int k = 0; k < o->sendProcessCount; ++k)
        if (o->sendProcessID[k] == r) po = k;
      if (po < 0) abort();
      int cnt = s->sendNeighbourCountCum[p + 1] - s->sendNeighbourCountCum[p];
      if (cnt != o->sendNeighbourCountCum[po + 1] - o->sendNeighbourCountCum[po])
        abort();
      const double *recv = &sendbuf[q][NDIM * o->sendNeighbourCountCum[po]];
      for (int i = 0; i < cnt; ++i) {
        int nodeIndex = NDIM * s->sendNodeIndex[s->sendNeighbourCountCum[p] + i];
        for (int l = 0; l < NDIM; ++l) a[nodeIndex + l] += recv[NDIM * i + l];
      }
    }
  }
  for (int r = 0; r < nranks; ++r) free(sendbuf[r]);
  free(sendbuf);
}

/* src/fem/SolidMechanics/CheckEnergy.cpp:19-52 */
void oracle_CheckEnergy_local(const oracle_state *s, double out[3]) {
  double sum_Wint_n = 0.0, sum_Wext_n = 0.0, delta_d = 0.0, WKE = 0.0;
  /* ownership scan (:21-33), done once instead of per node: O(shared) */
  char *skip = (char *)calloc(s->nNodes > 0 ? s->nNodes : 1, 1);
  for (int j = 0; j < s->sendProcessCount; ++j)
    if (s->sendProcessID[j] < s->world_rank)
      for (int k = s->sendNeighbourCountCum[j]; k < s->sendNeighbourCountCum[j + 1]; ++k)
        skip[s->sendNodeIndex[k]] = 1;
  for (int i = 0; i < s->nNodes; ++i) {
    if (skip[i]) continue;
    int index = i * NDIM;
    for (int j = 0; j < NDIM; ++j) {
      int indexJ = index + j;
      delta_d = s->displacements[indexJ] - s->displacements_prev[indexJ];
      WKE += s->mass[indexJ] * s->velocities[indexJ] * s->velocities[indexJ];
      if (s->boundary[indexJ]) {
        double reaction = s->fi_prev[indexJ] + s->fi[indexJ] +
                          s->mass[indexJ] * (s->accelerations[indexJ] +
                                             s->accelerations_prev[indexJ]);
        sum_Wext_n += delta_d * reaction;
      }
      sum_Wint_n += delta_d * (s->fi_prev[indexJ] + s->fi[indexJ]);
      sum_Wext_n += delta_d * (s->fe_prev[indexJ] + s->fe[indexJ]);
    }
  }
  free(skip);
  WKE *= 0.5;
  sum_Wint_n *= 0.5;
  sum_Wext_n *= 0.5;
  out[0] = WKE;
  out[1] = sum_Wint_n;
  out[2] = sum_Wext_n;
}

/* examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:184-244 as a
 * descriptor (see header) */
static void applyBC(oracle_state *s, const int *bc_kind, const double *bc_rate) {
  const int nDOF = NDIM * s->nNodes;
  for (int i = 0; i < nDOF; ++i) {
    int k = bc_kind[i];
    if (k > 0) {
      s->boundary[i] = 1;
      s->displacements[i] = s->Time * bc_rate[k];
      s->velocities[i] = bc_rate[k];
      s->accelerations[i] = 0.0;
    }
  }
}

static double stable_dt_all(oracle_state **ranks, int nranks) {
  double dtMin = huge_dt;
  for (int r = 0; r < nranks; ++r) {
    double d = oracle_StableTimeStep_local(ranks[r]);
    if (d < dtMin) dtMin = d; /* MPI_Allreduce(MIN), StableTimeStep.cpp:33 */
  }
  return dtMin;
}

static int get_force_all(oracle_state **ranks, int nranks) {
  int bad = 0;
  for (int r = 0; r < nranks; ++r) bad |= oracle_GetForce_local(ranks[r]);
  if (nranks > 1) oracle_halo_sum(ranks, nranks, 0);
  for (int r = 0; r < nranks; ++r) oracle_GetForce_finish(ranks[r]);
  return bad;
}

/* examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:83-171 */
int oracle_run_explicit(oracle_state **ranks, int nranks, int *const *bc_kind,
                        const double *bc_rate, double tMax, int maxSteps,
                        double ExplicitTimeStepReduction,
                        double FailureTimeStep, int first_call,
                        double *dt_hist, double *energy_hist) {
  return oracle_run_explicit_injury(ranks, nranks, bc_kind, bc_rate, tMax, maxSteps, ExplicitTimeStepReduction,
                                    FailureTimeStep, first_call, dt_hist, energy_hist, NULL);
}

int oracle_run_explicit_injury(oracle_state **ranks, int nranks, int *const *bc_kind, const double *bc_rate,
                               double tMax, int maxSteps, double ExplicitTimeStepReduction, double FailureTimeStep,
                               int first_call, double *dt_hist, double *energy_hist, oracle_injury **inj) {
  double Time = ranks[0]->Time, dt = ranks[0]->dt;
  if (first_call) {
    for (int r = 0; r < nranks; ++r) applyBC(ranks[r], bc_kind[r], bc_rate);
    double dtMin = stable_dt_all(ranks, nranks);
    if (dtMin < FailureTimeStep) return -19;
    dt = ExplicitTimeStepReduction * dtMin;
    for (int r = 0; r < nranks; ++r) ranks[r]->dt = dt;
    if (get_force_all(ranks, nranks)) return -1;
    for (int r = 0; r < nranks; ++r) oracle_CalculateAccelerations(ranks[r]);
  }
  int steps = 0;
  while (Time < tMax && steps < maxSteps) {
    double t_n = Time;
    double t_np1 = Time + dt;
    Time = t_np1;
    double dt_nphalf = dt;
    double t_nphalf = 0.5 * (t_np1 + t_n);
    if (dt_hist) dt_hist[steps] = dt;
    for (int r = 0; r < nranks; ++r) {
      oracle_state *s = ranks[r];
      const int nDOF = NDIM * s->nNodes;
      s->Time = Time;
      for (int i = 0; i < nDOF; i++) {
        if (s->boundary[i]) {
          s->velocities_half[i] = s->velocities[i];
        } else {
          s->velocities_half[i] =
              s->velocities[i] + (t_nphalf - t_n) * s->accelerations[i];
        }
      }
      memcpy(s->displacements_prev, s->displacements, nDOF * sizeof(double));
      memcpy(s->accelerations_prev, s->accelerations, nDOF * sizeof(double));
      memcpy(s->fi_prev, s->fi, nDOF * sizeof(double));
      memcpy(s->fe_prev, s->fe, nDOF * sizeof(double));
      for (int i = 0; i < nDOF; i++)
        if (!s->boundary[i])
          s->displacements[i] = s->displacements[i] + dt_nphalf * s->velocities_half[i];
      applyBC(s, bc_kind[r], bc_rate);
    }
    if (get_force_all(ranks, nranks)) return -1;
    for (int r = 0; r < nranks; ++r) {
      oracle_state *s = ranks[r];
      const int nDOF = NDIM * s->nNodes;
      oracle_CalculateAccelerations(s);
      for (int i = 0; i < nDOF; i++)
        if (!s->boundary[i])
          s->velocities[i] =
              s->velocities_half[i] + (t_np1 - t_nphalf) * s->accelerations[i];
    }
    /* CheckEnergy.cpp:54-64: three reductions to rank 0, running sums */
    {
      double WKE_Total = 0.0, Wint_n_total = 0.0, Wext_n_total = 0.0;
      for (int r = 0; r < nranks; ++r) {
        double part[3];
        oracle_CheckEnergy_local(ranks[r], part);
        WKE_Total += part[0];
        Wint_n_total += part[1];
        Wext_n_total += part[2];
      }
      ranks[0]->Wint_n += Wint_n_total;
      ranks[0]->Wext_n += Wext_n_total;
      if (energy_hist) {
        energy_hist[4 * steps + 0] = ranks[0]->Wint_n;
        energy_hist[4 * steps + 1] = ranks[0]->Wext_n;
        energy_hist[4 * steps + 2] = WKE_Total;
        energy_hist[4 * steps + 3] =
            fabs(WKE_Total + ranks[0]->Wint_n - ranks[0]->Wext_n);
      }
    }
    if (inj) oracle_CalculateInjuryCriterions(ranks, inj, nranks, Time, dt); /* ex5.cpp:240 */
    steps++;
    double dtMin = stable_dt_all(ranks, nranks);
    if (dtMin < FailureTimeStep) {
      for (int r = 0; r < nranks; ++r) { ranks[r]->Time = Time; ranks[r]->dt = dt; }
      return -19;
    }
    dt = ExplicitTimeStepReduction * dtMin;
    for (int r = 0; r < nranks; ++r) ranks[r]->dt = dt;
  }
  for (int r = 0; r < nranks; ++r) { ranks[r]->Time = Time; ranks[r]->dt = dt; }
  return steps;
}

/* src/elements/ElementCalculations/CalculateStrain.cpp:77-97: dgemm with
 * alpha = 0.5/8, beta = 1 accumulating into E, then subtract 0.5 on the
 * diagonal */
void oracle_CalculateStrain(const oracle_state *s, double *Eavg) {
  for (int elm = 0; elm < s->nElements; ++elm) {
    double *E = &Eavg[9 * elm];
    for (int i = 0; i < 9; ++i) E[i] = 0.0;
    const int countGP = NGP(s, elm);
    double preFactor = 0.5 / ((double)countGP);
    for (int gp = 0; gp < countGP; ++gp) {
      const double *Fg = &s->F[9 * (GP0(s, elm) + gp)];
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
          double sum = 0.0;
          for (int l = 0; l < 3; ++l) sum += Fg[l + 3 * i] * Fg[l + 3 * j];
          E[i + 3 * j] = 1.0 * E[i + 3 * j] + preFactor * sum;
        }
    }
    E[0] -= 0.5;
    E[4] -= 0.5;
    E[8] -= 0.5;
  }
}

/* ------------------------------------------------------------------------ */
/* injury criteria (examples/ex5/ex5.cpp, src/elements/ElementCalculations/CalculateStrain.cpp, src/math/math.cpp) */

/* CalculateStrain.cpp:8-75 */
void oracle_CalculateMaximumPrincipalStrain(const oracle_state *s, int elm, double *Eavg_e, double *currentStrainMax,
                                            double *currentStrainMin, double *currentShearMax) {
  const double kPi = 4.0 * atan(1.0); /* :6 */
  double Eloc[9];
  double *E = Eavg_e ? Eavg_e : Eloc;
  for (int i = 0; i < 9; ++i) E[i] = 0.0;
  const int countGP = NGP(s, elm);
  double preFactor = 0.5 / ((double)countGP);
  for (int gp = 0; gp < countGP; ++gp) {
    const double *Fg = &s->F[9 * (GP0(s, elm) + gp)];
    for (int j = 0; j < 3; ++j) /* dgemm('T','N'), alpha = preFactor, beta = 1 */
      for (int i = 0; i < 3; ++i) {
        double sum = 0.0;
        for (int l = 0; l < 3; ++l) sum += Fg[l + 3 * i] * Fg[l + 3 * j];
        E[i + 3 * j] = 1.0 * E[i + 3 * j] + preFactor * sum;
      }
  }
  E[0] -= 0.5;
  E[4] -= 0.5;
  E[8] -= 0.5;
  double a = E[0], b = E[1], c = E[2], d = E[4], e = E[5], f = E[8];
  double p1 = b * b + c * c + e * e;
  double eps1, eps2, eps3;
  if (p1 == 0) {
    eps1 = a;
    eps2 = d;
    eps3 = f;
  } else {
    double I1 = a + d + f;
    double I2 = a * (d + f) + d * f - b * b - c * c - e * e;
    double I3 = a * d * f + 2.0 * b * c * e - b * b * f - c * c * d - e * e * a;
    double Q = (3.0 * I2 - I1 * I1) / 9.0;
    double R = (2.0 * I1 * I1 * I1 - 9.0 * I1 * I2 + 27.0 * I3) / 54.0;
    double theta = acos(R / sqrt(-Q * Q * Q));
    double sqrtQ = 2.0 * sqrt(-Q);
    I1 = I1 / 3.0;
    eps1 = sqrtQ * cos(theta / 3.0) + I1;
    eps2 = sqrtQ * cos((theta + 2.0 * kPi) / 3.0) + I1;
    eps3 = sqrtQ * cos((theta + 4.0 * kPi) / 3.0) + I1;
  }
  double min, max;
  max = fmax(eps3, fmax(eps2, eps1));
  min = fmin(eps3, fmin(eps2, eps1));
  *currentShearMax = 0.5 * (max - min);
  if (max > 0.0) *currentStrainMax = max; else *currentStrainMax = 0.0;
  if (min < 0.0) *currentStrainMin = min; else *currentStrainMin = 0.0;
}

/* ex5.cpp:1251-1306 */
void oracle_InitInjuryCriterion(const oracle_state *s, oracle_injury *inj, const int *injuryExcludePID,
                                int injuryExcludePIDCount) {
  int count = 0;
  for (int i = 0; i < s->nElements; i++) {
    int include = 1;
    int elementPID = s->pid[i];
    for (int j = 0; j < injuryExcludePIDCount; ++j)
      if (elementPID == injuryExcludePID[j]) { include = 0; break; }
    if (include) {
      inj->elementIDInjury[count] = i;
      count += 1;
    }
  }
  inj->nElementsInjury = count;
  for (int j = 0; j < count; ++j) {
    inj->MPSgt15[j] = inj->MPSgt30[j] = inj->MPSRgt120[j] = inj->MPSxSRgt28[j] = 0;
    inj->PS_Old[j] = 0.0;
    inj->PSxSRArray[j] = 0.0;
  }
  inj->maxStrain = inj->minStrain = inj->maxShear = inj->maxPSxSR = 0.0; /* ex5.cpp:62,73 */
  inj->maxElem = inj->minElem = inj->shearElem = inj->maxElemPSxSR = 0;  /* :63,74 */
  inj->maxT = inj->minT = inj->maxShearT = inj->maxTimePSxSR = 0.0;
  inj->maxMPS95 = inj->maxTimeMPS95 = inj->maxMPSxSR95 = inj->maxTimeMPSxSR95 = 0.0;
  inj->maxElemCountMPS95 = inj->maxElemCountMPSxSR95 = 0;
}

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x < y) ? -1 : (x > y) ? 1 : 0;
}

/* math.cpp:160-199: gather, nth_element at (int)(total*0.95)-1 */
double oracle_compute95thPercentileValue(double *const *data, const int *sizes, int nranks) {
  int totalSize = 0;
  for (int r = 0; r < nranks; ++r) totalSize += sizes[r];
  double *full = (double *)malloc((totalSize > 0 ? totalSize : 1) * sizeof(double));
  int o = 0;
  for (int r = 0; r < nranks; ++r) {
    memcpy(full + o, data[r], sizes[r] * sizeof(double));
    o += sizes[r];
  }
  int index95 = (int)(totalSize * 0.95) - 1; /* the reference faults when this is negative */
  double v = 0.0;
  if (index95 >= 0) {
    qsort(full, totalSize, sizeof(double), cmp_double);
    v = full[index95];
  }
  free(full);
  return v;
}

/* ex5.cpp:1311-1430 */
void oracle_CalculateInjuryCriterions(oracle_state **ranks, oracle_injury **injs, int nranks, double Time, double dt) {
  for (int r = 0; r < nranks; ++r) {
    const oracle_state *s = ranks[r];
    oracle_injury *q = injs[r];
    double currentStrainMaxElem, currentStrainMinElem, currentShearMaxElem;
    double PSR = 0.0, PSxSR = 0.0;
    for (int j = 0; j < q->nElementsInjury; j++) {
      int i = q->elementIDInjury[j];
      oracle_CalculateMaximumPrincipalStrain(s, i, NULL, &currentStrainMaxElem, &currentStrainMinElem,
                                             &currentShearMaxElem);
      if (q->maxStrain < currentStrainMaxElem) { q->maxStrain = currentStrainMaxElem; q->maxElem = i; q->maxT = Time; }
      if (q->minStrain > currentStrainMinElem) { q->minStrain = currentStrainMinElem; q->minElem = i; q->minT = Time; }
      if (q->maxShear < currentShearMaxElem) { q->maxShear = currentShearMaxElem; q->shearElem = i; q->maxShearT = Time; }
      if (!q->MPSgt15[j]) if (currentStrainMaxElem > 0.15) q->MPSgt15[j] = 1;
      if (!q->MPSgt30[j]) if (currentStrainMaxElem > 0.30) q->MPSgt30[j] = 1;
      PSR = (currentStrainMaxElem - q->PS_Old[j]) / dt;
      PSxSR = currentStrainMaxElem * PSR;
      if (q->maxPSxSR < PSxSR) { q->maxPSxSR = PSxSR; q->maxElemPSxSR = i; q->maxTimePSxSR = Time; }
      if (!q->MPSRgt120[j]) if (PSR > 120.0) q->MPSRgt120[j] = 1;
      if (!q->MPSxSRgt28[j]) if (PSxSR > 28.0) q->MPSxSRgt28[j] = 1;
      q->PS_Old[j] = currentStrainMaxElem;
      q->PSxSRArray[j] = PSxSR;
    }
  }
  double **arr = (double **)malloc(nranks * sizeof(double *));
  int *sz = (int *)malloc(nranks * sizeof(int));
  for (int pass = 0; pass < 2; ++pass) {
    for (int r = 0; r < nranks; ++r) {
      arr[r] = pass ? injs[r]->PSxSRArray : injs[r]->PS_Old;
      sz[r] = injs[r]->nElementsInjury;
    }
    double v95 = oracle_compute95thPercentileValue(arr, sz, nranks);
    for (int r = 0; r < nranks; ++r) {
      oracle_injury *q = injs[r];
      double *mx = pass ? &q->maxMPSxSR95 : &q->maxMPS95;
      double *mt = pass ? &q->maxTimeMPSxSR95 : &q->maxTimeMPS95;
      int *list = pass ? q->maxElemListMPSxSR95 : q->maxElemListMPS95;
      int *cnt = pass ? &q->maxElemCountMPSxSR95 : &q->maxElemCountMPS95;
      if (v95 > *mx) {
        *mx = v95;
        *mt = Time;
        int count = 0;
        for (int j = 0; j < q->nElementsInjury; j++)
          if (arr[r][j] >= *mx) { list[count] = q->elementIDInjury[j]; count = count + 1; }
        *cnt = count;
      }
    }
  }
  free(arr);
  free(sz);
}

/* ex5.cpp:1043-1066 with Elements.cpp:30-38 / CalculateCentroidAndVolume.cpp:26-37 */
void oracle_injury_volumes(const oracle_state *s, const oracle_injury *inj, double out[5]) {
  for (int k = 0; k < 5; ++k) out[k] = 0.0;
  for (int j = 0; j < inj->nElementsInjury; ++j) {
    int e = inj->elementIDInjury[j];
    double coord[24];
    for (int a = 0; a < NEN(s, e); ++a)
      for (int k = 0; k < 3; ++k) coord[3 * a + k] = s->coordinates[3 * s->connectivity[C0(s, e) + a] + k];
    double eV = NEN(s, e) == 8 ? oracle_volumeHexahedron(coord) : volumeTetrahedron(coord);
    if (inj->MPSgt15[j]) {
      out[0] += eV;
      if (inj->MPSgt30[j]) out[1] += eV;
    }
    if (inj->MPSRgt120[j]) out[2] += eV;
    if (inj->MPSxSRgt28[j]) out[3] += eV;
    out[4] += eV;
  }
}

/* ------------------------------------------------------------------------ */
/* rigid-body prescribed motion (examples/ex5/ex5.cpp, src/math/math.cpp) */

/* math.cpp:99-119 (the out-of-range error message is not reproduced) */
double oracle_interpolateLinear(int n, const double *x, const double *y, double value) {
  if (value < x[0]) return 0.0;
  if (value > x[n - 1]) return y[n - 1];
  if (value == x[0]) return y[0];
  int index = 0;
  for (int i = 1; i < n; ++i) {
    if (value <= x[i]) { index = i - 1; break; }
  }
  const double yValue = y[index] + (y[index + 1] - y[index]) * (value - x[index]) / (x[index + 1] - x[index]);
  return yValue;
}
/* math.cpp:50-54 */
static void crossProduct(const double *a, const double *b, double *result) {
  result[0] = a[1] * b[2] - a[2] * b[1];
  result[1] = -a[0] * b[2] + a[2] * b[0];
  result[2] = a[0] * b[1] - a[1] * b[0];
}
/* math.cpp:122-132 */
void oracle_quaternionExp(const double *q1, double *q2) {
  double vMag = q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3];
  if (vMag == 0) {
    q2[0] = 1.0; q2[1] = 0.0; q2[2] = 0.0; q2[3] = 0.0;
  } else {
    vMag = sqrt(vMag);
    const double d1 = exp(q1[0]);
    const double d2 = d1 * sin(vMag) / vMag;
    q2[0] = d1 * cos(vMag); q2[1] = d2 * q1[1]; q2[2] = d2 * q1[2]; q2[3] = d2 * q1[3];
  }
}
/* math.cpp:134-139 */
static void quaternionMultiply(const double *q1, const double *q2, double *qr) {
  qr[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
  qr[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
  qr[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
  qr[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
}
/* math.cpp:141-145 */
static void quaternionInverse(const double *q, double *qinv) {
  double norm = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  qinv[0] = q[0] / norm; qinv[1] = -q[1] / norm;
  qinv[2] = -q[2] / norm; qinv[3] = -q[3] / norm;
}
/* math.cpp:154-158 */
void oracle_quaternionRotate(const double *v, const double *R, const double *Rinv, double *vp) {
  double Rv[4];
  quaternionMultiply(R, v, Rv);
  quaternionMultiply(Rv, Rinv, vp);
}

/* ex5.cpp:976-1020 */
void oracle_computeDerivatives(const oracle_rigid *rb, const double *y, double *ydot, double t) {
  ydot[0] = oracle_interpolateLinear(rb->size[0], rb->t[0], rb->v[0], t);
  ydot[1] = oracle_interpolateLinear(rb->size[1], rb->t[1], rb->v[1], t);
  ydot[2] = oracle_interpolateLinear(rb->size[2], rb->t[2], rb->v[2], t);
  ydot[6] = oracle_interpolateLinear(rb->size[3], rb->t[3], rb->v[3], t);
  ydot[7] = oracle_interpolateLinear(rb->size[4], rb->t[4], rb->v[4], t);
  ydot[8] = oracle_interpolateLinear(rb->size[5], rb->t[5], rb->v[5], t);
  ydot[9] = y[6];
  ydot[10] = y[7];
  ydot[11] = y[8];
  double r[3];
  r[0] = y[3]; r[1] = y[4]; r[2] = y[5];
  double *rdot = &(ydot[3]);
  double rMagnitude = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (rMagnitude < 1e-10) {
    rdot[0] = 0.5 * y[0];
    rdot[1] = 0.5 * y[1];
    rdot[2] = 0.5 * y[2];
  } else {
    double rCotR = rMagnitude / tan(rMagnitude);
    double omega[3]; omega[0] = y[0]; omega[1] = y[1]; omega[2] = y[2];
    crossProduct(omega, r, rdot);
    for (int i = 0; i < 3; ++i) r[i] = r[i] / rMagnitude;
    double rDotOmega = r[0] * omega[0] + r[1] * omega[1] + r[2] * omega[2];
    for (int i = 0; i < 3; ++i) rdot[i] = 0.5 * (rdot[i] + rCotR * y[i] + (1.0 - rCotR) * rDotOmega * r[i]);
  }
}

/* boost/numeric/odeint/stepper/runge_kutta_dopri5.hpp, do_step_impl(system, in, dxdt_in, t, out, dxdt_out, dt) with
 * in == out, dxdt_in == dxdt_out (what do_step(sys, x, dxdt, t, dt) does for an FSAL stepper).  Published
 * Dormand-Prince coefficients; sums evaluated left to right as odeint's scale_sumN functors do.  UNPINNED (no Boost
 * in this image). */
void oracle_dopri5_step(const oracle_rigid *rb, double *x, double *dxdt, double t, double dt) {
  const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
  const double b21 = 1.0 / 5.0;
  const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
  const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
  const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
  const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
               b65 = -5103.0 / 18656.0;
  const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
  double xt[12], k2[12], k3[12], k4[12], k5[12], k6[12];
  const double *k1 = dxdt;
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b21 * k1[i];
  oracle_computeDerivatives(rb, xt, k2, t + dt * a2);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b31 * k1[i] + dt * b32 * k2[i];
  oracle_computeDerivatives(rb, xt, k3, t + dt * a3);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b41 * k1[i] + dt * b42 * k2[i] + dt * b43 * k3[i];
  oracle_computeDerivatives(rb, xt, k4, t + dt * a4);
  for (int i = 0; i < 12; ++i)
    xt[i] = 1.0 * x[i] + dt * b51 * k1[i] + dt * b52 * k2[i] + dt * b53 * k3[i] + dt * b54 * k4[i];
  oracle_computeDerivatives(rb, xt, k5, t + dt * a5);
  for (int i = 0; i < 12; ++i)
    xt[i] = 1.0 * x[i] + dt * b61 * k1[i] + dt * b62 * k2[i] + dt * b63 * k3[i] + dt * b64 * k4[i] + dt * b65 * k5[i];
  oracle_computeDerivatives(rb, xt, k6, t + dt);
  for (int i = 0; i < 12; ++i)
    xt[i] = 1.0 * x[i] + dt * c1 * k1[i] + dt * c3 * k3[i] + dt * c4 * k4[i] + dt * c5 * k5[i] + dt * c6 * k6[i];
  for (int i = 0; i < 12; ++i) x[i] = xt[i];
  oracle_computeDerivatives(rb, x, dxdt, t + dt);
}

static int cmp_int(const void *a, const void *b) { return (*(const int *)a - *(const int *)b); }

/* ex5.cpp:819-911 */
void oracle_InitRigidBoundary(oracle_state *s, oracle_rigid *rb) {
  int rigidNodeCount = 0;
  for (int i = 0; i < s->nElements; ++i)
    if (s->materialID[s->pid[i]] == 0) rigidNodeCount += NEN(s, i);
  int *rigidNodeID = (int *)malloc((rigidNodeCount > 0 ? rigidNodeCount : 1) * sizeof(int));
  int nodePtr = 0;
  for (int i = 0; i < s->nElements; ++i)
    if (s->materialID[s->pid[i]] == 0)
      for (int j = C0(s, i); j < C0(s, i) + NEN(s, i); ++j) rigidNodeID[nodePtr++] = s->connectivity[j];
  qsort(rigidNodeID, rigidNodeCount, sizeof(int), cmp_int);
  for (int i = 0; i < rigidNodeCount; ++i) {
    int index = rigidNodeID[i] * NDIM;
    s->boundary[index] = 1;
    s->boundary[index + 1] = 1;
    s->boundary[index + 2] = 1;
  }
  free(rigidNodeID);
  int idIndex = 0;
  for (int i = 0; i < s->nNodes; ++i)
    if (s->boundary[i * NDIM]) rb->boundaryID[idIndex++] = i;
  rb->boundarySize = idIndex;
  for (int j = 0; j < 12; ++j) { rb->y[j] = 0.0; rb->ydot[j] = 0.0; }
  for (int i = 0; i < rb->boundarySize; i++) {
    int index = rb->boundaryID[i] * NDIM;
    for (int j = 0; j < NDIM; ++j) {
      s->displacements[index + j] = 0.0;
      s->velocities[index + j] = 0.0;
      s->accelerations[index + j] = 0.0;
    }
  }
}

/* ex5.cpp:339-371 */
void oracle_ApplyAccBoundaryConditions(oracle_state *s, oracle_rigid *rb, double Time, double dt) {
  double r[4], R[4], Rinv[4], V[4], Vp[4];
  double omegaR[3], omega[3], omegaOmegaR[3], omegaVel[3], vel[3];
  double alpha[3], alphaR[3], locV[3];
  double *yInt = rb->y, *ydotInt = rb->ydot;
  oracle_dopri5_step(rb, yInt, ydotInt, Time - dt, dt);
  r[0] = 0.0; r[1] = yInt[3]; r[2] = yInt[4]; r[3] = yInt[5];
  oracle_quaternionExp(r, R);
  quaternionInverse(R, Rinv);
  omega[0] = yInt[0]; omega[1] = yInt[1]; omega[2] = yInt[2];
  alpha[0] = ydotInt[0]; alpha[1] = ydotInt[1]; alpha[2] = ydotInt[2];
  vel[0] = yInt[6]; vel[1] = yInt[7]; vel[2] = yInt[8];
  for (int i = 0; i < rb->boundarySize; i++) {
    int index = rb->boundaryID[i] * NDIM;
    for (int j = 0; j < NDIM; ++j) locV[j] = s->coordinates[index + j];
    V[0] = 0.0; V[1] = locV[0]; V[2] = locV[1]; V[3] = locV[2];
    oracle_quaternionRotate(V, R, Rinv, Vp);
    crossProduct(omega, &(Vp[1]), omegaR);
    crossProduct(omega, omegaR, omegaOmegaR);
    crossProduct(omega, vel, omegaVel);
    crossProduct(alpha, &(Vp[1]), alphaR);
    for (int j = 0; j < NDIM; ++j) {
      s->displacements[index + j] = Vp[j + 1] - locV[j] + yInt[9 + j];
      s->velocities[index + j] = omegaR[j] + yInt[6 + j];
      s->accelerations[index + j] = 2.0 * omegaVel[j] + omegaOmegaR[j] + ydotInt[6 + j] + alphaR[j];
    }
  }
}

/* ex5.cpp:159-295 */
int oracle_run_explicit_rigid(oracle_state *s, oracle_rigid *rb, double tMax, int maxSteps,
                              double ExplicitTimeStepReduction, double FailureTimeStep, int first_call, double *dt_hist,
                              double *energy_hist, oracle_injury *inj) {
  double Time = s->Time, dt = s->dt;
  const int nDOF = NDIM * s->nNodes;
  if (first_call) {
    double dtMin = oracle_StableTimeStep_local(s);
    if (dtMin < FailureTimeStep) return -19;
    dt = ExplicitTimeStepReduction * dtMin;
    s->dt = dt;
    if (oracle_GetForce_local(s)) return -1;
    oracle_GetForce_finish(s);
    oracle_CalculateAccelerations(s);
  }
  int steps = 0;
  while (Time < tMax && steps < maxSteps) {
    double t_n = Time;
    double t_np1 = Time + dt;
    Time = t_np1;
    double dt_nphalf = dt;
    double t_nphalf = 0.5 * (t_np1 + t_n);
    if (dt_hist) dt_hist[steps] = dt;
    s->Time = Time;
    for (int i = 0; i < nDOF; i++) {
      if (s->boundary[i]) s->velocities_half[i] = s->velocities[i];
      else s->velocities_half[i] = s->velocities[i] + (t_nphalf - t_n) * s->accelerations[i];
    }
    memcpy(s->displacements_prev, s->displacements, nDOF * sizeof(double));
    memcpy(s->accelerations_prev, s->accelerations, nDOF * sizeof(double));
    memcpy(s->fi_prev, s->fi, nDOF * sizeof(double));
    memcpy(s->fe_prev, s->fe, nDOF * sizeof(double));
    for (int i = 0; i < nDOF; i++)
      if (!s->boundary[i]) s->displacements[i] = s->displacements[i] + dt_nphalf * s->velocities_half[i];
    oracle_ApplyAccBoundaryConditions(s, rb, Time, dt);
    if (oracle_GetForce_local(s)) return -1;
    oracle_GetForce_finish(s);
    oracle_CalculateAccelerations(s);
    for (int i = 0; i < nDOF; i++)
      if (!s->boundary[i]) s->velocities[i] = s->velocities_half[i] + (t_np1 - t_nphalf) * s->accelerations[i];
    {
      double part[3];
      oracle_CheckEnergy_local(s, part);
      s->Wint_n += part[1];
      s->Wext_n += part[2];
      if (energy_hist) {
        energy_hist[4 * steps + 0] = s->Wint_n;
        energy_hist[4 * steps + 1] = s->Wext_n;
        energy_hist[4 * steps + 2] = part[0];
        energy_hist[4 * steps + 3] = fabs(part[0] + s->Wint_n - s->Wext_n);
      }
    }
    if (inj) {
      oracle_state *rs[1] = {s};
      oracle_injury *ri[1] = {inj};
      oracle_CalculateInjuryCriterions(rs, ri, 1, Time, dt);
    }
    steps++;
    double dtMin = oracle_StableTimeStep_local(s);
    if (dtMin < FailureTimeStep) { s->Time = Time; s->dt = dt; return -19; }
    dt = ExplicitTimeStepReduction * dtMin;
    s->dt = dt;
  }
  s->Time = Time;
  s->dt = dt;
  return steps;
}
