// elem_harness.cpp -- TEST ONLY.  Compiles femtech_b200/csrc/hex8_element.cuh
// (the exact arithmetic the CUDA kernels run, one element per thread) with g++
// so the `-m "not gpu"` suite can compare it with the oracle without a GPU.
// Never loaded by the product.
#include "../../femtech_b200/csrc/hex8_element.cuh"

namespace {
struct HostHist {
  double* h;  // [8][18]: h1[6], h2[6], s0[6] per Gauss point
  void load(int gp, ftb::GpHistory& g) const {
    for (int i = 0; i < 6; ++i) { g.h1[i] = h[18 * gp + i]; g.h2[i] = h[18 * gp + 6 + i]; g.s0[i] = h[18 * gp + 12 + i]; }
  }
  void store(int gp, const ftb::GpHistory& g) const {
    for (int i = 0; i < 6; ++i) { h[18 * gp + i] = g.h1[i]; h[18 * gp + 6 + i] = g.h2[i]; h[18 * gp + 12 + i] = g.s0[i]; }
  }
};
struct HostOut {
  static constexpr bool enabled = true;
  static constexpr bool want_S = true;
  double *F, *detF, *pk2;
  void put(int gp, const double Fm[3][3], double J, const double Sv[6]) const {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) F[9 * gp + 3 * j + i] = Fm[i][j];  // reference layout: column-major
    detF[gp] = J;
    for (int i = 0; i < 6; ++i) pk2[6 * gp + i] = Sv[i];
  }
};
}  // namespace

extern "C" int harness_element(const double* X24, const double* U24, int mat, const double* mp, double* hist144,
                               int updHist, double* fe24, double* dtElem, double* F72, double* detF8, double* pk2_48) {
  double X[8][3], U[8][3], fe[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) { X[k][c] = X24[3 * k + c]; U[k][c] = U24[3 * k + c]; }
  HostHist hh{hist144};
  HostOut ho{F72, detF8, pk2_48};
  ftb::LocalScratch sc;
  int st = ftb::hex8_element<-1, true>(X, U, mat, mp, updHist != 0, hh, ho, sc, fe, dtElem);
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) fe24[3 * k + c] = fe[k][c];
  return st;
}

extern "C" double harness_mass(const double* X24, double rho, double* me8) {
  double X[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) X[k][c] = X24[3 * k + c];
  return ftb::hex8_lumped_mass(X, rho, me8);
}

// the affine-reference-geometry variant of the hexahedron (parallelepipeds); returns -1 if the element does not qualify
extern "C" int harness_element_affine(const double* X24, const double* U24, int mat, const double* mp, double* hist144,
                                      int updHist, double* fe24, double* dtElem, double* F72, double* detF8, double* pk2_48) {
  double X[8][3], U[8][3], fe[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) { X[k][c] = X24[3 * k + c]; U[k][c] = U24[3 * k + c]; }
  if (!ftb::hex8_is_affine(X)) return -1;
  HostHist hh{hist144};
  HostOut ho{F72, detF8, pk2_48};
  ftb::LocalScratchAffine sc;
  int st = ftb::hex8_element_affine_in<-1, true>(ftb::ArrayInAffine{X, U}, mat, mp, updHist != 0, hh, ho, sc, fe, dtElem);
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) fe24[3 * k + c] = fe[k][c];
  return st;
}

// The slot plan of k_elem_affine's asynchronous gather, replayed on the host: components 1 and 2 of the nodal input wait
// in the scratch slots FTB_ASTAGE_U / FTB_ASTAGE_X that the element function later overwrites with columns and cofactors.
// Any aliasing between a staged value and an earlier store would change the result.
namespace {
struct HostStagedAffine {
  const double* x0;  // component 0 of nodes 0, 1, 3, 4
  const double* u0;  // component 0 of the 8 nodes
  const double* v;   // the scratch itself
  void getX(const int c, double x[4]) const {
    for (int k = 0; k < 4; ++k) x[k] = (c == 0) ? x0[k] : v[FTB_ASTAGE_X(k, c)];
  }
  void getU(const int c, double nu[8]) const {
    for (int k = 0; k < 8; ++k) nu[k] = (c == 0) ? u0[k] : v[FTB_ASTAGE_U(k, c)];
  }
};
}  // namespace
extern "C" int harness_element_affine_staged(const double* X24, const double* U24, int mat, const double* mp, double* hist144,
                                             int updHist, double* fe24, double* dtElem, double* F72, double* detF8, double* pk2_48) {
  double X[8][3], fe[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) X[k][c] = X24[3 * k + c];
  if (!ftb::hex8_is_affine(X)) return -1;
  const int nx[4] = {0, 1, 3, 4};
  ftb::LocalScratchAffine sc;
  for (int i = 0; i < FTB_AFFINE_SLOTS; ++i) sc.v[i] = 1e300;  // poison
  double x0[4], u0[8];
  for (int k = 0; k < 8; ++k) u0[k] = U24[3 * k];
  for (int k = 0; k < 4; ++k) x0[k] = X24[3 * nx[k]];
  for (int c = 1; c < 3; ++c) {
    for (int k = 0; k < 8; ++k) sc.v[FTB_ASTAGE_U(k, c)] = U24[3 * k + c];
    for (int k = 0; k < 4; ++k) sc.v[FTB_ASTAGE_X(k, c)] = X24[3 * nx[k] + c];
  }
  HostHist hh{hist144};
  HostOut ho{F72, detF8, pk2_48};
  int st = ftb::hex8_element_affine_in<-1, true>(HostStagedAffine{x0, u0, sc.v}, mat, mp, updHist != 0, hh, ho, sc, fe, dtElem);
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) fe24[3 * k + c] = fe[k][c];
  return st;
}

// hex8_element_affine_cj (neo-Hookean / HGO parallelepiped, current-Jacobian form: what k_elem_affine_cj runs), direct and through
// the staging-slot plan of the kernel (components 1, 2 of the input wait in slots the function later overwrites)
template <int MAT>
static int affine_cj(const double* X24, const double* U24, const double* mp, int staged, double* fe24, double* dtElem) {
  double X[8][3], U[8][3], fe[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) { X[k][c] = X24[3 * k + c]; U[k][c] = U24[3 * k + c]; }
  if (!ftb::hex8_is_affine(X)) return -1;
  ftb::LocalScratchAffine sc;
  for (int i = 0; i < FTB_AFFINE_SLOTS; ++i) sc.v[i] = 1e300;  // poison
  int st;
  if (staged) {
    const int nx[4] = {0, 1, 3, 4};
    double x0[4], u0[8];
    for (int k = 0; k < 8; ++k) u0[k] = U24[3 * k];
    for (int k = 0; k < 4; ++k) x0[k] = X24[3 * nx[k]];
    for (int c = 1; c < 3; ++c) {
      for (int k = 0; k < 8; ++k) sc.v[FTB_ASTAGE_U(k, c)] = U24[3 * k + c];
      for (int k = 0; k < 4; ++k) sc.v[FTB_ASTAGE_X(k, c)] = X24[3 * nx[k] + c];
    }
    st = ftb::hex8_element_affine_cj<MAT, true>(HostStagedAffine{x0, u0, sc.v}, mp, sc, fe, dtElem);
    for (int i = FTB_NH_SLOTS; i < FTB_AFFINE_SLOTS; ++i)
      if (sc.v[i] != 1e300) return -2;  // the function must stay inside its 44 slots
  } else {
    st = ftb::hex8_element_affine_cj<MAT, true>(ftb::ArrayInAffine{X, U}, mp, sc, fe, dtElem);
  }
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) fe24[3 * k + c] = fe[k][c];
  return st;
}

extern "C" int harness_element_affine_cj(const double* X24, const double* U24, int mat, const double* mp, int staged, double* fe24, double* dtElem) {
  if (mat == 1) return affine_cj<1>(X24, U24, mp, staged, fe24, dtElem);
  if (mat == 4) return affine_cj<4>(X24, U24, mp, staged, fe24, dtElem);
  return -3;
}

// hex8_brick_setup + hex8_brick_loop (the element function of k_brick, cut where the kernel's pipeline cuts it): the
// reference nodes 0, 1, 3, 4 and the displacements come from node-indexed tables like the kernel's shared-memory staging
namespace {
struct HostStagedBrick {
  const double* xtab;    // [3][4] component-major reference coordinates of nodes 0, 1, 3, 4
  const double* utab;    // [3][8] component-major displacement table
  void getX(const int c, double x[4]) const {
    for (int k = 0; k < 4; ++k) x[k] = xtab[4 * c + k];
  }
  void getU(const int c, double nu[8]) const {
    for (int k = 0; k < 8; ++k) nu[k] = utab[8 * c + k];
  }
};
}  // namespace
extern "C" int harness_element_brick(const double* X24, const double* U24, int mat, const double* mp, double* hist144,
                                     int updHist, double* fe24, double* dtElem, double* F72, double* detF8, double* pk2_48) {
  double X[8][3], fe[8][3];
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) X[k][c] = X24[3 * k + c];
  if (!ftb::hex8_is_affine(X)) return -1;
  const int nx[4] = {0, 1, 3, 4};
  ftb::LocalScratchBrick sc;
  for (int i = 0; i < 36; ++i) sc.col[i] = 1e300;  // poison
  for (int i = 0; i < 9; ++i) sc.ji[i] = 1e300;
  double utab[24], xtab[12];
  for (int c = 0; c < 3; ++c) {
    for (int k = 0; k < 8; ++k) utab[8 * c + k] = U24[3 * k + c];
    for (int k = 0; k < 4; ++k) xtab[4 * c + k] = X24[3 * nx[k] + c];
  }
  HostHist hh{hist144};
  HostOut ho{F72, detF8, pk2_48};
  double det, dtk;
  int st = ftb::hex8_brick_setup(HostStagedBrick{xtab, utab}, mp, sc, &det, &dtk);
  st |= ftb::hex8_brick_loop<-1>(mat, mp, updHist != 0, hh, ho, sc, det, dtk, fe, dtElem);
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) fe24[3 * k + c] = fe[k][c];
  return st;
}

// largest face area of a hexahedron given the 8 CURRENT nodal positions: the filtered form the kernels use and the
// every-face form it replaces
extern "C" void harness_face_amax(const double* x24, double* out2) {
  double xm[7][3], n[8], g[7];
  for (int c = 0; c < 3; ++c) {
    for (int k = 0; k < 8; ++k) n[k] = x24[3 * k + c];
    ftb::hex_modes(n, g);
    for (int m = 0; m < 7; ++m) xm[m][c] = g[m];
  }
  out2[0] = ftb::hex_face_amax(xm);
  out2[1] = ftb::hex_face_amax_all(xm);
}

// CalculateMaximumPrincipalStrain of one element: out = max, min, shear, then the 6 sums of F^T F
extern "C" void harness_principal(const double* X24, const double* U24, double* out9) {
  double X[8][3], U[8][3], fe[8][3], d;
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c) { X[k][c] = X24[3 * k + c]; U[k][c] = U24[3 * k + c]; }
  double mp[ftb::FTB_MP_STRIDE] = {0};
  double cs[6] = {0, 0, 0, 0, 0, 0};
  ftb::LocalScratch sc;
  ftb::hex8_element<0, false>(X, U, 0, mp, false, ftb::NoHistory(), ftb::StrainSink{cs}, sc, fe, &d);
  ftb::principal_strains(cs, &out9[0], &out9[1], &out9[2]);
  for (int i = 0; i < 6; ++i) out9[3 + i] = cs[i];
}

// one C3D4 element: the arithmetic of the mixed-mesh branch of k_elem
extern "C" int harness_tet(const double* X12, const double* U12, int mat, const double* mp, double* hist18, int updHist,
                           double* fe12, double* dtElem, double* F9, double* detF1, double* pk2_6, double* me4) {
  double X[4][3], U[4][3], fe[4][3];
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < 3; ++c) { X[k][c] = X12[3 * k + c]; U[k][c] = U12[3 * k + c]; }
  HostHist hh{hist18};
  HostOut ho{F9, detF1, pk2_6};
  int st = ftb::tet4_element<-1, true>(X, U, mat, mp, updHist != 0, hh, ho, fe, dtElem);
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < 3; ++c) fe12[3 * k + c] = fe[k][c];
  if (me4) ftb::tet4_lumped_mass(X, mp[ftb::MP_RHO], me4);
  return st;
}
