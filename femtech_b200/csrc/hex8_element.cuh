// hex8_element.cuh -- per-element fp64 arithmetic of the FemTech hex8 explicit
// step, written for one CUDA thread per element (sm_100a).
//
// What it replaces (all citations relative to /root/reference):
//   GetForce_3D element/GP loop            src/fem/SolidMechanics/GetForce_3D.cpp:15-46
//   CalculateDeformationGradient           src/elements/ElementCalculations/CalculateDeformationGradient.cpp:4-30
//   DeterminateF / InverseF                src/math/DeterminateF.cpp:38-57, src/math/InverseF.cpp:38-66
//   StressUpdate + src/materials/*.cpp     src/fem/SolidMechanics/StressUpdate.cpp:5-29
//   InternalForceUpdate (B^T sigma)        src/fem/SolidMechanics/InternalForceUpdate.cpp:4-28
//   CalculateTimeStep / char. length       src/timestep/CalculateTimeStep.cpp:7-21,
//                                          src/elements/CharacteristicLength/CalculateCharacteristicLength_C3D8.cpp:3-30,
//                                          src/math/Geometry.cpp:3-64
//   ShapeFunction_C3D8 (dN/dX, detJ)       src/fem/ShapeFunctions/ShapeFunction_C3D8.cpp:4-128
//
// This is NOT a transcription.  The reference stores dN/dX (192 doubles per
// element), builds a 6x24 B matrix per Gauss point and calls dgemv.  Here
// nothing per-Gauss-point is stored; the trilinear hex is handled in its
// Walsh-Hadamard ("mode") basis:
//   x(xi,eta,zeta) = 1/8 [ G0 + xi g1 + eta g2 + zeta g3 + xi.eta g12
//                          + eta.zeta g23 + xi.zeta g13 + xi.eta.zeta g123 ]
// so the Jacobian at a Gauss point (+-a,+-a,+-a) is a signed sum of 4 mode
// vectors per column, F = I + (dU/dxi)(dX/dxi)^-1, the nodal forces are
//   f_k = sum_gp  P cof(J0) grad_xi N_k   with  P = F S  (first Piola-Kirchhoff)
// accumulated in the same mode basis (7 force modes, the 8th is zero by
// momentum balance) and transformed back to the 8 nodes by one butterfly.
// Mathematically identical to w detJ0 B^T S; floating-point results differ from
// the reference at the 1e-15 level per step (tests pin <= 1e-9 after 1000 steps).
//
// Everything here is __host__ __device__ so that tests/ can compile the same
// arithmetic with g++ and compare it with the oracle WITHOUT a GPU.  The
// product never runs it on the host.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FTB_HD __host__ __device__ __forceinline__
#else
#define FTB_HD inline
#endif

namespace ftb {

// Gauss abscissa exactly as written in the reference (15 digits, not 1/sqrt(3)):
// src/fem/ShapeFunctions/GaussQuadrature3D.cpp:19-49
#define FTB_GP_A 0.577350269189626

FTB_HD double ftb_rcbrt(const double x) {
#if defined(__CUDA_ARCH__)
  return rcbrt(x);
#else
  return 1.0 / cbrt(x);
#endif
}

// ln J and 1/J of the neo-Hookean stress, and the reciprocal of the HGO stresses, without the library's special-case
// handling.  One Gauss point of k_elem_affine spends 172 fp64 and ~125 other instructions, and the kernel is bound by the
// issue port: `log(J)` and `1.0 / J` alone were ~35 fp64 and ~45 other instructions of those
// (exponent extraction, polynomial coefficients moved through uniform registers, slow-path tests and calls).  Here:
//   1/x    MUFU.RCP64H seed + one cubic correction step (x is a positive normal number: J > 0 is checked by the caller);
//   ln J   = 2 atanh(s), s = (J - 1)/(J + 1), as 2 s (1 + z (1/3 + z (1/5 + ... + z/27))), z = s^2, for |s| <= 1/4
//          (0.6 <= J <= 1.667; truncation < 1e-17 relative), coefficients as constant-bank operands of the DFMAs;
//          outside that range the library functions.
// Agreement with log / division: a few ulp (the parity tests hold the kernels to 1e-9 against the oracle over 1000 steps).
#if defined(__CUDACC__)
__constant__ double FTB_LN_C[13] = {1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0,  1.0 / 9.0,  1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0,
                                    1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0, 1.0 / 25.0, 1.0 / 27.0};
#endif
FTB_HD double ftb_rcp(const double x) {
#if defined(__CUDA_ARCH__) && !defined(FTB_LIBM_MATERIAL)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // one cubic step instead of two Newton steps: 1/x = r (1 + e + e^2 + ...), e = 1 - x r; the seed is good to ~2^-20
  // (MUFU.RCP64H reads the upper word of x), so dropping e^3 leaves 2^-60 relative
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  return 1.0 / x;
#endif
}
FTB_HD void ftb_ln_rcp(const double J, double* lnJ, double* rJ) {
#if defined(__CUDA_ARCH__) && !defined(FTB_LIBM_MATERIAL)
  const double w = J - 1.0, t = J + 1.0;
  const double rt = ftb_rcp(t);
  double s = w * rt;
  s = fma(rt, fma(-s, t, w), s);  // one correction: s = (J - 1)/(J + 1) to within an ulp
  *rJ = ftb_rcp(J);
  if (fabs(s) <= 0.25) {
    // even and odd coefficients as two Horner chains in z^2 (7 + 6 dependent FMAs instead of 13: the chain is what a
    // warp waits for here)
    const double z = s * s, zz = z * z;
    double pe = FTB_LN_C[12], po = FTB_LN_C[11];
#pragma unroll
    for (int k = 10; k >= 0; k -= 2) pe = fma(pe, zz, FTB_LN_C[k]);
#pragma unroll
    for (int k = 9; k >= 1; k -= 2) po = fma(po, zz, FTB_LN_C[k]);
    const double p = fma(po, z, pe);
    const double s2 = s + s;
    *lnJ = fma(s2 * z, p, s2);
  } else {
    *lnJ = log(J);
  }
#else
  *lnJ = log(J);
  *rJ = 1.0 / J;
#endif
}

// Per-part parameter block (device memory, FTB_MP_STRIDE doubles per part).
// [0..8] are properties[9*pid + k] (src/io/input/ReadMaterials.cpp:43-122).
enum {
  MP_RHO = 0, MP_MU = 1, MP_LAMBDA = 2, MP_K1 = 3, MP_K2 = 4, MP_G1 = 5, MP_T1 = 6, MP_G2 = 7, MP_T2 = 8,
  MP_CE = 9,    // dilatational wave speed, CalculateTimeStep.cpp:15-18 (host, same formula)
  MP_KBULK = 10, // lambda + 2 mu / 3, HGOIsotropic.cpp:44
  MP_C11 = 11, MP_C12 = 12, MP_C21 = 13, MP_C22 = 14, // Prony factors of the current dt
  MP_MATID = 15,
  FTB_MP_STRIDE = 16
};

// Mode index: 0:g1 1:g2 2:g3 3:g12 4:g23 5:g13 6:g123
// Node k of C3D8 has signs (ShapeFunction_C3D8.cpp:22-29):
//   0(---) 1(+--) 2(++-) 3(-+-) 4(--+) 5(+-+) 6(+++) 7(-++)
// With b = bit0(xi+) | bit1(eta+) | bit2(zeta+) the nodes in binary order are
//   v[0]=n0 v[1]=n1 v[2]=n3 v[3]=n2 v[4]=n4 v[5]=n5 v[6]=n7 v[7]=n6.

// Forward Walsh-Hadamard transform of 8 nodal scalars -> 7 modes (the constant
// mode is not needed).  n[] in C3D8 node order.
FTB_HD void hex_modes(const double n[8], double g[7]) {
  // stage xi
  const double a0 = n[1] + n[0], d0 = n[1] - n[0];  // (eta-,zeta-)
  const double a1 = n[2] + n[3], d1 = n[2] - n[3];  // (eta+,zeta-)
  const double a2 = n[5] + n[4], d2 = n[5] - n[4];  // (eta-,zeta+)
  const double a3 = n[6] + n[7], d3 = n[6] - n[7];  // (eta+,zeta+)
  // stage eta
  const double aa0 = a1 + a0, ad0 = a1 - a0;  // zeta-
  const double aa1 = a3 + a2, ad1 = a3 - a2;  // zeta+
  const double da0 = d1 + d0, dd0 = d1 - d0;
  const double da1 = d3 + d2, dd1 = d3 - d2;
  // stage zeta
  g[0] = da1 + da0;  // s1
  g[1] = ad1 + ad0;  // s2
  g[2] = aa1 - aa0;  // s3
  g[3] = dd1 + dd0;  // s1 s2
  g[4] = ad1 - ad0;  // s2 s3
  g[5] = da1 - da0;  // s1 s3
  g[6] = dd1 - dd0;  // s1 s2 s3
}

// Inverse: nodal values f_k = s1 p0 + s2 p1 + s3 p2 + s1s2 p3 + s2s3 p4 + s1s3 p5 + s1s2s3 p6
FTB_HD void hex_modes_to_nodes(const double p[7], double f[8]) {
  // combine per (eta,zeta) sign pair: A = terms without s1, B = terms with s1
  // f = A(s2,s3) + s1 * B(s2,s3)
  const double A_mm = -p[1] - p[2] + p[4];          // s2=-,s3=-
  const double A_pm = p[1] - p[2] - p[4];           // s2=+,s3=-
  const double A_mp = -p[1] + p[2] - p[4];          // s2=-,s3=+
  const double A_pp = p[1] + p[2] + p[4];           // s2=+,s3=+
  const double B_mm = p[0] - p[3] - p[5] + p[6];
  const double B_pm = p[0] + p[3] - p[5] - p[6];
  const double B_mp = p[0] - p[3] + p[5] - p[6];
  const double B_pp = p[0] + p[3] + p[5] + p[6];
  f[0] = A_mm - B_mm;
  f[1] = A_mm + B_mm;
  f[2] = A_pm + B_pm;
  f[3] = A_pm - B_pm;
  f[4] = A_mp - B_mp;
  f[5] = A_mp + B_mp;
  f[6] = A_pp + B_pp;
  f[7] = A_pp - B_pp;
}

// Gauss-point signs in the reference's numbering (GaussQuadrature3D.cpp:19-49)
#define FTB_GP_S1(gp) (((gp) == 1 || (gp) == 2 || (gp) == 5 || (gp) == 6) ? 1.0 : -1.0)
#define FTB_GP_S2(gp) (((gp) == 2 || (gp) == 3 || (gp) == 6 || (gp) == 7) ? 1.0 : -1.0)
#define FTB_GP_S3(gp) (((gp) < 4) ? 1.0 : -1.0)

// 8*dX/dxi at a Gauss point from pre-scaled modes m[7][3] (m[3..5] already
// multiplied by a, m[6] by a^2).  J[i][c]: i = space component, c = xi,eta,zeta.
FTB_HD void gp_jacobian(const double m[7][3], const double s1, const double s2, const double s3, double J[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    J[i][0] = m[0][i] + s2 * m[3][i] + s3 * m[5][i] + (s2 * s3) * m[6][i];
    J[i][1] = m[1][i] + s1 * m[3][i] + s3 * m[4][i] + (s1 * s3) * m[6][i];
    J[i][2] = m[2][i] + s2 * m[4][i] + s1 * m[5][i] + (s1 * s2) * m[6][i];
  }
}

// cofactor matrix: cof[i][j] = cofactor of A[i][j];  A^-1 = cof^T / det
FTB_HD void cofactor3(const double A[3][3], double C[3][3]) {
  C[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  C[0][1] = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  C[0][2] = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  C[1][0] = A[0][2] * A[2][1] - A[0][1] * A[2][2];
  C[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0];
  C[1][2] = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  C[2][0] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
  C[2][1] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
  C[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
}

// Material models ---------------------------------------------------------
// Input: F (row-major F[i][J]), cofF = cof(F), J = det F, part parameters mp.
// Output: first Piola-Kirchhoff stress P = F S used by the force contraction,
// and (when wantS) the PK2 stress in the reference's Voigt order
// [11,22,33,23,13,12] for output parity.
// hist: for material 5, pointers to the 6-component history of this Gauss
// point (H1,H2,S0n each [6], Voigt order); updated in place when updHist.
struct GpHistory {
  double h1[6], h2[6], s0[6];
};

// symmetric 3x3 in Voigt order [11,22,33,23,13,12] -> P = F * S
FTB_HD void F_times_symS(const double F[3][3], const double S[6], double P[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    P[i][0] = F[i][0] * S[0] + F[i][1] * S[5] + F[i][2] * S[4];
    P[i][1] = F[i][0] * S[5] + F[i][1] * S[1] + F[i][2] * S[3];
    P[i][2] = F[i][0] * S[4] + F[i][1] * S[3] + F[i][2] * S[2];
  }
}

// S = scale * cof^T * sig * cof for symmetric sig (Voigt), result Voigt.
FTB_HD void pullback_sym(const double cofF[3][3], const double sig[6], const double scale, double S[6]) {
  double T[3][3];  // T = sig * cof
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    T[0][j] = sig[0] * cofF[0][j] + sig[5] * cofF[1][j] + sig[4] * cofF[2][j];
    T[1][j] = sig[5] * cofF[0][j] + sig[1] * cofF[1][j] + sig[3] * cofF[2][j];
    T[2][j] = sig[4] * cofF[0][j] + sig[3] * cofF[1][j] + sig[2] * cofF[2][j];
  }
  S[0] = scale * (cofF[0][0] * T[0][0] + cofF[1][0] * T[1][0] + cofF[2][0] * T[2][0]);
  S[1] = scale * (cofF[0][1] * T[0][1] + cofF[1][1] * T[1][1] + cofF[2][1] * T[2][1]);
  S[2] = scale * (cofF[0][2] * T[0][2] + cofF[1][2] * T[1][2] + cofF[2][2] * T[2][2]);
  S[3] = scale * (cofF[0][1] * T[0][2] + cofF[1][1] * T[1][2] + cofF[2][1] * T[2][2]);
  S[4] = scale * (cofF[0][0] * T[0][2] + cofF[1][0] * T[1][2] + cofF[2][0] * T[2][2]);
  S[5] = scale * (cofF[0][0] * T[0][1] + cofF[1][0] * T[1][1] + cofF[2][0] * T[2][1]);
}

// Cauchy stress of the HGO model with isotropic fibre dispersion
// (src/materials/HGOIsotropic.cpp:44-84): sigma = pref*dev(B) + hydro*I.
FTB_HD void hgo_cauchy(const double F[3][3], const double J, const double* __restrict__ mp, double sig[6]) {
  const double mu = mp[MP_MU], k1 = mp[MP_K1], k2 = mp[MP_K2], K = mp[MP_KBULK];
  const double rJ = ftb_rcp(J);
  const double hydro = 0.5 * K * (J * J - 1.0) * rJ;
  double B[6];  // B = F F^T, Voigt
  B[0] = F[0][0] * F[0][0] + F[0][1] * F[0][1] + F[0][2] * F[0][2];
  B[1] = F[1][0] * F[1][0] + F[1][1] * F[1][1] + F[1][2] * F[1][2];
  B[2] = F[2][0] * F[2][0] + F[2][1] * F[2][1] + F[2][2] * F[2][2];
  B[3] = F[1][0] * F[2][0] + F[1][1] * F[2][1] + F[1][2] * F[2][2];
  B[4] = F[0][0] * F[2][0] + F[0][1] * F[2][1] + F[0][2] * F[2][2];
  B[5] = F[0][0] * F[1][0] + F[0][1] * F[1][1] + F[0][2] * F[1][2];
  const double trB = B[0] + B[1] + B[2];
  const double rc = ftb_rcbrt(J);
  const double Jm23 = rc * rc;  // pow(J, -2/3)
  const double I1 = Jm23 * trB;
  const double kappa = 1.0 / 3.0;
  const double Ea = kappa * (I1 - 3.0);
  double fiber = 0.0;
  if (Ea > 0.0) {
    const double ex = (k2 == 0.0) ? 1.0 : exp(k2 * Ea * Ea);
    fiber = 2.0 * k1 * ex * Ea * kappa;
  }
  const double pref = Jm23 * (mu + fiber) * rJ;
  const double t3 = trB * (1.0 / 3.0);
  sig[0] = (B[0] - t3) * pref + hydro;
  sig[1] = (B[1] - t3) * pref + hydro;
  sig[2] = (B[2] - t3) * pref + hydro;
  sig[3] = B[3] * pref;
  sig[4] = B[4] * pref;
  sig[5] = B[5] * pref;
}

// Returns 0, or 1 for an unknown material id (StressUpdate.cpp:24-26).
template <bool WANT_S>
FTB_HD int material_P(const int mat, const double F[3][3], const double cofF[3][3], const double J,
                      const double* __restrict__ mp, GpHistory* hist, const bool updHist, double P[3][3],
                      double Sv[6]) {
  const double mu = mp[MP_MU], lambda = mp[MP_LAMBDA];
  switch (mat) {
    case 0: {  // rigid part: pk2 stays 0 (StressUpdate.cpp:8-9)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[i][j] = 0.0;
      if (WANT_S)
#pragma unroll
        for (int i = 0; i < 6; ++i) Sv[i] = 0.0;
      return 0;
    }
    case 1: {  // compressible neo-Hookean (CompressibleNeoHookean.cpp:43-55)
      // S = mu (I - C^-1) + lambda ln J C^-1  =>  P = mu F + (lambda ln J - mu) F^-T,  F^-T = cof F / J
      // the reciprocal runs concurrently with the logarithm (two independent serial chains)
      double lnJ, rJ;
      ftb_ln_rcp(J, &lnJ, &rJ);
      const double c = (lambda * lnJ - mu) * rJ;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[i][j] = mu * F[i][j] + c * cofF[i][j];
      if (WANT_S) {
        // C^-1 = F^-1 F^-T = cof^T cof / J^2
        const double r = 1.0 / (J * J), l = lambda * log(J);
        double Ci[6];
        Ci[0] = r * (cofF[0][0] * cofF[0][0] + cofF[1][0] * cofF[1][0] + cofF[2][0] * cofF[2][0]);
        Ci[1] = r * (cofF[0][1] * cofF[0][1] + cofF[1][1] * cofF[1][1] + cofF[2][1] * cofF[2][1]);
        Ci[2] = r * (cofF[0][2] * cofF[0][2] + cofF[1][2] * cofF[1][2] + cofF[2][2] * cofF[2][2]);
        Ci[3] = r * (cofF[0][1] * cofF[0][2] + cofF[1][1] * cofF[1][2] + cofF[2][1] * cofF[2][2]);
        Ci[4] = r * (cofF[0][0] * cofF[0][2] + cofF[1][0] * cofF[1][2] + cofF[2][0] * cofF[2][2]);
        Ci[5] = r * (cofF[0][0] * cofF[0][1] + cofF[1][0] * cofF[1][1] + cofF[2][0] * cofF[2][1]);
        Sv[0] = mu * (1.0 - Ci[0]) + l * Ci[0];
        Sv[1] = mu * (1.0 - Ci[1]) + l * Ci[1];
        Sv[2] = mu * (1.0 - Ci[2]) + l * Ci[2];
        Sv[3] = -mu * Ci[3] + l * Ci[3];
        Sv[4] = -mu * Ci[4] + l * Ci[4];
        Sv[5] = -mu * Ci[5] + l * Ci[5];
      }
      return 0;
    }
    case 2: {  // St Venant-Kirchhoff (StVenantKirchhoff.cpp:25-40)
      double E[6], S[6];
      E[0] = 0.5 * (F[0][0] * F[0][0] + F[1][0] * F[1][0] + F[2][0] * F[2][0]) - 0.5;
      E[1] = 0.5 * (F[0][1] * F[0][1] + F[1][1] * F[1][1] + F[2][1] * F[2][1]) - 0.5;
      E[2] = 0.5 * (F[0][2] * F[0][2] + F[1][2] * F[1][2] + F[2][2] * F[2][2]) - 0.5;
      E[3] = 0.5 * (F[0][1] * F[0][2] + F[1][1] * F[1][2] + F[2][1] * F[2][2]);
      E[4] = 0.5 * (F[0][0] * F[0][2] + F[1][0] * F[1][2] + F[2][0] * F[2][2]);
      E[5] = 0.5 * (F[0][0] * F[0][1] + F[1][0] * F[1][1] + F[2][0] * F[2][1]);
      const double trE = E[0] + E[1] + E[2];
#pragma unroll
      for (int i = 0; i < 6; ++i) S[i] = 2.0 * mu * E[i];
      S[0] += lambda * trE;
      S[1] += lambda * trE;
      S[2] += lambda * trE;
      F_times_symS(F, S, P);
      if (WANT_S)
#pragma unroll
        for (int i = 0; i < 6; ++i) Sv[i] = S[i];
      return 0;
    }
    case 3: {  // "linear elastic" (LinearElastic.cpp:30-63): eps = F + F^T - 2I (i.e. 2 eps),
      // Pm = mu eps + lambda (tr F - 3) I, S = F^-1 Pm.  The reference keeps only the entries
      // S11,S22,S33,S23,S13,S12 of the (non-symmetric) product and uses them as a symmetric tensor.
      const double trEps = F[0][0] + F[1][1] + F[2][2] - 3.0;
      double Pm[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Pm[i][j] = mu * ((F[i][j] + F[j][i]) - (i == j ? 2.0 : 0.0));
      Pm[0][0] += lambda * trEps;
      Pm[1][1] += lambda * trEps;
      Pm[2][2] += lambda * trEps;
      // F^-1[a][b] = cofF[b][a] / J
      double S[6];
      const double rJ = 1.0 / J;
      S[0] = rJ * (cofF[0][0] * Pm[0][0] + cofF[1][0] * Pm[1][0] + cofF[2][0] * Pm[2][0]);
      S[1] = rJ * (cofF[0][1] * Pm[0][1] + cofF[1][1] * Pm[1][1] + cofF[2][1] * Pm[2][1]);
      S[2] = rJ * (cofF[0][2] * Pm[0][2] + cofF[1][2] * Pm[1][2] + cofF[2][2] * Pm[2][2]);
      S[3] = rJ * (cofF[0][1] * Pm[0][2] + cofF[1][1] * Pm[1][2] + cofF[2][1] * Pm[2][2]);  // S23
      S[4] = rJ * (cofF[0][0] * Pm[0][2] + cofF[1][0] * Pm[1][2] + cofF[2][0] * Pm[2][2]);  // S13
      S[5] = rJ * (cofF[0][0] * Pm[0][1] + cofF[1][0] * Pm[1][1] + cofF[2][0] * Pm[2][1]);  // S12
      F_times_symS(F, S, P);
      if (WANT_S)
#pragma unroll
        for (int i = 0; i < 6; ++i) Sv[i] = S[i];
      return 0;
    }
    case 4: {  // HGO, isotropic fibre dispersion (HGOIsotropic.cpp:21-115)
      double sig[6];
      hgo_cauchy(F, J, mp, sig);
      // S = J F^-1 sig F^-T ;  P = F S = J sig F^-T = sig * cof F
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        P[0][j] = sig[0] * cofF[0][j] + sig[5] * cofF[1][j] + sig[4] * cofF[2][j];
        P[1][j] = sig[5] * cofF[0][j] + sig[1] * cofF[1][j] + sig[3] * cofF[2][j];
        P[2][j] = sig[4] * cofF[0][j] + sig[3] * cofF[1][j] + sig[2] * cofF[2][j];
      }
      if (WANT_S) pullback_sym(cofF, sig, 1.0 / J, Sv);
      return 0;
    }
    case 5: {  // HGO + 2-term Prony viscoelasticity (HGOIsotropicViscoelastic.cpp:27-168)
      double sig[6], S[6];
      hgo_cauchy(F, J, mp, sig);
      const double rJ = 1.0 / J;
      pullback_sym(cofF, sig, rJ, S);  // S0 = J F^-1 sig F^-T = cof^T sig cof / J
      // C = F^T F ; C:S
      double Cm[6];
      Cm[0] = F[0][0] * F[0][0] + F[1][0] * F[1][0] + F[2][0] * F[2][0];
      Cm[1] = F[0][1] * F[0][1] + F[1][1] * F[1][1] + F[2][1] * F[2][1];
      Cm[2] = F[0][2] * F[0][2] + F[1][2] * F[1][2] + F[2][2] * F[2][2];
      Cm[3] = F[0][1] * F[0][2] + F[1][1] * F[1][2] + F[2][1] * F[2][2];
      Cm[4] = F[0][0] * F[0][2] + F[1][0] * F[1][2] + F[2][0] * F[2][2];
      Cm[5] = F[0][0] * F[0][1] + F[1][0] * F[1][1] + F[2][0] * F[2][1];
      double SddC = Cm[0] * S[0] + Cm[1] * S[1] + Cm[2] * S[2] + 2.0 * (Cm[3] * S[3] + Cm[4] * S[4] + Cm[5] * S[5]);
      SddC = SddC / 3.0;
      // Sic = SddC * F^-1 F^-T = SddC * cof^T cof / J^2
      const double sc = SddC * rJ * rJ;
      double Sdev[6];
      Sdev[0] = S[0] - sc * (cofF[0][0] * cofF[0][0] + cofF[1][0] * cofF[1][0] + cofF[2][0] * cofF[2][0]);
      Sdev[1] = S[1] - sc * (cofF[0][1] * cofF[0][1] + cofF[1][1] * cofF[1][1] + cofF[2][1] * cofF[2][1]);
      Sdev[2] = S[2] - sc * (cofF[0][2] * cofF[0][2] + cofF[1][2] * cofF[1][2] + cofF[2][2] * cofF[2][2]);
      Sdev[3] = S[3] - sc * (cofF[0][1] * cofF[0][2] + cofF[1][1] * cofF[1][2] + cofF[2][1] * cofF[2][2]);
      Sdev[4] = S[4] - sc * (cofF[0][0] * cofF[0][2] + cofF[1][0] * cofF[1][2] + cofF[2][0] * cofF[2][2]);
      Sdev[5] = S[5] - sc * (cofF[0][0] * cofF[0][1] + cofF[1][0] * cofF[1][1] + cofF[2][0] * cofF[2][1]);
      if (updHist) {
        const double c11 = mp[MP_C11], c12 = mp[MP_C12], c21 = mp[MP_C21], c22 = mp[MP_C22];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double dS = Sdev[i] - hist->s0[i];
          hist->h1[i] = c11 * hist->h1[i] + c21 * dS;
          hist->h2[i] = c12 * hist->h2[i] + c22 * dS;
          hist->s0[i] = Sdev[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) S[i] = S[i] + hist->h1[i] + hist->h2[i];
      F_times_symS(F, S, P);
      if (WANT_S)
#pragma unroll
        for (int i = 0; i < 6; ++i) Sv[i] = S[i];
      return 0;
    }
    default:
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[i][j] = 0.0;
      return 1;
  }
}

// Stable time step of one element from the modes of the CURRENT coordinates
// xm[7][3] (unscaled modes of X+u).  src/math/Geometry.cpp:3-64 restated in
// the mode basis: the six +-sum vectors q0..q5 of volumeHexahedron are exactly
// g12,g13,g1,g23,g2,g3; a face's (centerD,c1,c2) are (mode +- mode)/8.
FTB_HD double tp3(const double s[3], const double a[3], const double b[3]) {  // math.cpp:44-48
  return s[2] * (a[0] * b[1] - a[1] * b[0]) + s[0] * (a[1] * b[2] - a[2] * b[1]) - s[1] * (a[0] * b[2] - a[2] * b[0]);
}
FTB_HD double ncross(const double a[3], const double b[3]) {  // math.cpp:30-41
  const double z = a[0] * b[1] - a[1] * b[0];
  const double x = a[1] * b[2] - a[2] * b[1];
  const double y = -a[0] * b[2] + a[2] * b[0];
  return sqrt(x * x + y * y + z * z);
}
// Faces come in opposite pairs (zeta=-+1, xi=-+1, eta=-+1).  For a pair, with unscaled (x8) vectors
//   centerD = mD -+ m123,  c1 = mA -+ mAn,  c2 = mB -+ mBn
// the cross products share their terms: c1 x c2 = (mA x mB + mAn x mBn) -+ (mA x mBn + mAn x mB).
// The parallelogram branch (Geometry.cpp:46-48, signed test kept) needs only |c1 x c2|^2, so the
// square root is taken once per element; the 2x2 Gauss branch (:50-63) is the rare slow path.
FTB_HD void cross3(const double a[3], const double b[3], double r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
FTB_HD double face_area_gauss(const double cD8[3], const double c18[3], const double c28[3]) {
  double cD[3], c1[3], c2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { cD[i] = 0.125 * cD8[i]; c1[i] = 0.125 * c18[i]; c2[i] = 0.125 * c28[i]; }
  const double t = sqrt(3.0) / 3.0;
  double area = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double qi = i ? t : -t, qj = j ? t : -t;
      double v1[3], v2[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        v1[k] = qj * cD[k] + c1[k];
        v2[k] = qi * cD[k] + c2[k];
      }
      area += ncross(v1, v2);
    }
  return area;
}
// updates n2max (largest |c1 x c2|^2 of the parallelogram faces, x8 vectors) and aslow (largest
// area among faces that took the Gauss branch)
FTB_HD void face_pair(const double mD[3], const double m123[3], const double mA[3], const double mAn[3],
                      const double mB[3], const double mBn[3], double& n2max, double& aslow) {
  double p1[3], p2[3], q1[3], q2[3];
  cross3(mA, mB, p1);
  cross3(mAn, mBn, p2);
  cross3(mA, mBn, q1);
  cross3(mAn, mB, q2);
  const double tol8 = 8.0e-6;  // centerD[i] < 1e-6  <=>  8 centerD[i] < 8e-6 (exact power-of-two scaling)
#pragma unroll
  for (int sgn = 0; sgn < 2; ++sgn) {
    const double sg = sgn ? 1.0 : -1.0;
    const double d0 = mD[0] + sg * m123[0], d1 = mD[1] + sg * m123[1], d2 = mD[2] + sg * m123[2];
    if ((d0 < tol8) && (d1 < tol8) && (d2 < tol8)) {
      const double x = (p1[0] + p2[0]) + sg * (q1[0] + q2[0]);
      const double y = (p1[1] + p2[1]) + sg * (q1[1] + q2[1]);
      const double z = (p1[2] + p2[2]) + sg * (q1[2] + q2[2]);
      const double n2 = x * x + y * y + z * z;
      if (n2 > n2max) n2max = n2;
    } else {
      double cD[3] = {d0, d1, d2}, c1[3], c2[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { c1[i] = mA[i] + sg * mAn[i]; c2[i] = mB[i] + sg * mBn[i]; }
      const double ar = face_area_gauss(cD, c1, c2);
      if (ar > aslow) aslow = ar;
    }
  }
}
// largest face area A_max of the element (CalculateCharacteristicLength_C3D8.cpp:10-27 over Geometry.cpp:29-64),
// every face evaluated in full: the straightforward form, kept as the checker of hex_face_amax below
FTB_HD double hex_face_amax_all(const double xm[7][3]) {
  const double* g1 = xm[0]; const double* g2 = xm[1]; const double* g3 = xm[2];
  const double* g12 = xm[3]; const double* g23 = xm[4]; const double* g13 = xm[5]; const double* g123 = xm[6];
  double n2max = 0.0, aslow = 0.0;
  face_pair(g12, g123, g1, g13, g2, g23, n2max, aslow);  // zeta = -+1: faces {0,1,2,3},{4,5,6,7}
  face_pair(g23, g123, g2, g12, g3, g13, n2max, aslow);  // xi   = -+1: faces {0,3,7,4},{1,2,6,5}
  face_pair(g13, g123, g1, g12, g3, g23, n2max, aslow);  // eta  = -+1: faces {0,1,5,4},{3,2,6,7}
  // parallelogram area = 4 |c1 x c2| with c = (x8 vector)/8  ->  |x8 cross| / 16
  const double afast = sqrt(n2max) * 0.0625;
  return afast > aslow ? afast : aslow;
}

// The same maximum with the square roots only where they can matter.  A face that takes the Gauss branch
// (Geometry.cpp:50-63) has the area  A = sum over the four points of |v1 x v2|,  v1 = c1 + qj cD, v2 = c2 + qi cD,
// qi, qj = +-1/sqrt(3).  With N0 = c1 x c2, N1 = c1 x cD, N2 = cD x c2 the four cross products are N0 + qi N1 + qj N2, so
//     4 |N0|  <=  A  <=  2 sqrt( sum |.|^2 ) = 4 sqrt( |N0|^2 + (|N1|^2 + |N2|^2)/3 )
// (triangle inequality / Cauchy-Schwarz; the mixed terms of the sum cancel over the four sign pairs).  Pass 1 computes
// both bounds for all six faces without a square root -- for a parallelogram face (:46-48) they coincide with the exact
// area -- and only the Gauss faces whose upper bound reaches the largest lower bound are then integrated (24 square roots
// per distorted element become 4, sometimes 8).  The result is the reference's maximum, formed from the same four
// norms per face; it differs from hex_face_amax_all by rounding only (tests pin 1e-13).
// Units: mode vectors are 8 x the reference's centerD/c1/c2, cross products 64 x; areas below are 64 x the true ones.
FTB_HD void face_pair_bounds(const double mD[3], const double m123[3], const double mA[3], const double mAn[3],
                             const double mB[3], const double mBn[3], const int f0, double L2[6], double U2[6], unsigned& gmask) {
  double p1[3], p2[3], q1[3], q2[3];
  cross3(mA, mB, p1);
  cross3(mAn, mBn, p2);
  cross3(mA, mBn, q1);
  cross3(mAn, mB, q2);
  const double tol8 = 8.0e-6;
  const double t = sqrt(3.0) / 3.0, t2 = t * t;
#pragma unroll
  for (int sgn = 0; sgn < 2; ++sgn) {
    const double sg = sgn ? 1.0 : -1.0;
    const double cD[3] = {mD[0] + sg * m123[0], mD[1] + sg * m123[1], mD[2] + sg * m123[2]};
    const double x = (p1[0] + p2[0]) + sg * (q1[0] + q2[0]);
    const double y = (p1[1] + p2[1]) + sg * (q1[1] + q2[1]);
    const double z = (p1[2] + p2[2]) + sg * (q1[2] + q2[2]);
    const double n0 = x * x + y * y + z * z;  // |c1 x c2|^2
    L2[f0 + sgn] = 16.0 * n0;
    if ((cD[0] < tol8) && (cD[1] < tol8) && (cD[2] < tol8)) {
      U2[f0 + sgn] = 16.0 * n0;
    } else {
      double c1[3], c2[3], N1[3], N2[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { c1[i] = mA[i] + sg * mAn[i]; c2[i] = mB[i] + sg * mBn[i]; }
      cross3(c1, cD, N1);
      cross3(cD, c2, N2);
      const double n12 = (N1[0] * N1[0] + N1[1] * N1[1] + N1[2] * N1[2]) + (N2[0] * N2[0] + N2[1] * N2[1] + N2[2] * N2[2]);
      U2[f0 + sgn] = 16.0 * (n0 + t2 * n12);
      gmask |= 1u << (f0 + sgn);
    }
  }
}
FTB_HD double sel3(const int p, const double a, const double b, const double c) { return p == 0 ? a : (p == 1 ? b : c); }
FTB_HD double hex_face_amax(const double xm[7][3]) {
  const double* g1 = xm[0]; const double* g2 = xm[1]; const double* g3 = xm[2];
  const double* g12 = xm[3]; const double* g23 = xm[4]; const double* g13 = xm[5]; const double* g123 = xm[6];
  double L2[6], U2[6];
  unsigned gmask = 0;  // bit f: face f takes the Gauss branch
  face_pair_bounds(g12, g123, g1, g13, g2, g23, 0, L2, U2, gmask);  // zeta = -+1
  face_pair_bounds(g23, g123, g2, g12, g3, g13, 2, L2, U2, gmask);  // xi   = -+1
  face_pair_bounds(g13, g123, g1, g12, g3, g23, 4, L2, U2, gmask);  // eta  = -+1
  double M = 0.0, para2 = 0.0;
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    M = L2[f] > M ? L2[f] : M;
    if (!((gmask >> f) & 1u)) para2 = L2[f] > para2 ? L2[f] : para2;
  }
  double best = sqrt(para2);  // exact for the parallelogram faces
  if (gmask) {
    unsigned cand = 0;
#pragma unroll
    for (int f = 0; f < 6; ++f)
      if (((gmask >> f) & 1u) && U2[f] * (1.0 + 1e-12) >= M) cand |= 1u << f;
    const double t = sqrt(3.0) / 3.0;
    while (cand) {  // each lane integrates ITS next candidate: one trip for most elements, whatever the faces are
      int f = 0;
#if defined(__CUDA_ARCH__)
      f = __ffs((int)cand) - 1;
#else
      while (!((cand >> f) & 1u)) ++f;
#endif
      cand &= cand - 1;
      const int p = f >> 1;
      const double sg = (f & 1) ? 1.0 : -1.0;
      double cD[3], c1[3], c2[3], N0[3], N1[3], N2[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        cD[i] = sel3(p, g12[i], g23[i], g13[i]) + sg * g123[i];
        c1[i] = sel3(p, g1[i], g2[i], g1[i]) + sg * sel3(p, g13[i], g12[i], g12[i]);
        c2[i] = sel3(p, g2[i], g3[i], g3[i]) + sg * sel3(p, g23[i], g13[i], g23[i]);
      }
      cross3(c1, c2, N0);
      cross3(c1, cD, N1);
      cross3(cD, c2, N2);
      double area = 0.0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double qi = i ? t : -t, qj = j ? t : -t;
          const double vx = N0[0] + qi * N1[0] + qj * N2[0];
          const double vy = N0[1] + qi * N1[1] + qj * N2[1];
          const double vz = N0[2] + qi * N1[2] + qj * N2[2];
          area += sqrt(vx * vx + vy * vy + vz * vz);
        }
      best = area > best ? area : best;
    }
  }
  return best * (1.0 / 64.0);
}
// returns V / A_max (CalculateCharacteristicLength_C3D8.cpp:3-30)
FTB_HD double hex_char_length(const double xm[7][3]) {
  const double* g1 = xm[0]; const double* g2 = xm[1]; const double* g3 = xm[2];
  const double* g12 = xm[3]; const double* g23 = xm[4]; const double* g13 = xm[5]; const double* g123 = xm[6];
  // volumeHexahedron: q0=g12 q1=g13 q2=g1 q3=g23 q4=g2 q5=g3
  const double vol = (tp3(g12, g2, g23) + tp3(g1, g12, g13) + tp3(g13, g23, g3)) / 192.0 + tp3(g1, g2, g3) / 64.0;
  double n2max = 0.0, aslow = 0.0;
  face_pair(g12, g123, g1, g13, g2, g23, n2max, aslow);  // zeta = -+1: faces {0,1,2,3},{4,5,6,7}
  face_pair(g23, g123, g2, g12, g3, g13, n2max, aslow);  // xi   = -+1: faces {0,3,7,4},{1,2,6,5}
  face_pair(g13, g123, g1, g12, g3, g23, n2max, aslow);  // eta  = -+1: faces {0,1,5,4},{3,2,6,7}
  // parallelogram area = 4 |c1 x c2| with c = (x8 vector)/8  ->  |x8 cross| / 16
  const double afast = sqrt(n2max) * 0.0625;
  const double amax = afast > aslow ? afast : aslow;
  return vol / amax;
}

// Element accessor for the Prony history: the kernel supplies load/store
// functors so that the same arithmetic runs on SoA device planes and on the
// host test harness.
struct NoHistory {
  FTB_HD void load(int, GpHistory&) const {}
  FTB_HD void store(int, const GpHistory&) const {}
};

// Scratch for the 2 x 12 Jacobian column vectors of an element (72 doubles).  The CUDA kernel
// keeps them in shared memory ([72][blockDim] doubles, conflict free) so that only the 21 force
// modes stay in registers across the Gauss-point loop; the host harness uses a plain array.
struct LocalScratch {
  double v[72];
  FTB_HD void st(int i, double x) { v[i] = x; }
  FTB_HD double ld(int i) const { return v[i]; }
};
// index of component c of column type t (0 xi, 1 eta, 2 zeta), sign combination q, field f (0 X, 1 U)
#define FTB_COL(f, t, q, c) ((((f) * 3 + (t)) * 4 + (q)) * 3 + (c))

// r(sa,sb) = A + sa B + sb C + sa sb D for the four sign pairs: q = (sa>0) + 2 (sb>0)
template <class Scratch>
FTB_HD void col_butterfly(Scratch& S, const int f, const int t, const int c, const double A, const double B,
                          const double C, const double D) {
  const double ad = A + D, am = A - D, bc = B + C, bm = B - C;
  S.st(FTB_COL(f, t, 3, c), ad + bc);  // (+,+)
  S.st(FTB_COL(f, t, 0, c), ad - bc);  // (-,-)
  S.st(FTB_COL(f, t, 1, c), am + bm);  // (+,-)
  S.st(FTB_COL(f, t, 2, c), am - bm);  // (-,+)
}

// Output sink for K_out (F, detF, pk2 in the reference's layouts); the hot
// kernel uses NoOutput.
struct NoOutput {
  static constexpr bool enabled = false;
  static constexpr bool want_S = false;
  FTB_HD void put(int, const double[3][3], double, const double[6]) const {}
};

// Sink of the injury criteria: sum over the Gauss points of F^T F (6 unique entries 00 11 22 12 02 01), the raw
// material of CalculateMaximumPrincipalStrain (CalculateStrain.cpp:16-24).  Lives in registers of the hot kernel.
struct StrainSink {
  static constexpr bool enabled = true;
  static constexpr bool want_S = false;
  double* c;
  FTB_HD void put(int, const double F[3][3], double, const double[6]) const {
    c[0] += F[0][0] * F[0][0] + F[1][0] * F[1][0] + F[2][0] * F[2][0];
    c[1] += F[0][1] * F[0][1] + F[1][1] * F[1][1] + F[2][1] * F[2][1];
    c[2] += F[0][2] * F[0][2] + F[1][2] * F[1][2] + F[2][2] * F[2][2];
    c[3] += F[0][1] * F[0][2] + F[1][1] * F[1][2] + F[2][1] * F[2][2];
    c[4] += F[0][0] * F[0][2] + F[1][0] * F[1][2] + F[2][0] * F[2][2];
    c[5] += F[0][0] * F[0][1] + F[1][0] * F[1][1] + F[2][0] * F[2][1];
  }
};

// CalculateStrain.cpp:8-75: E = (0.5/8) sum_gp F^T F - 0.5 I, closed-form eigenvalues of the symmetric 3x3,
// max clipped at >= 0, min at <= 0, shear = (max - min)/2 of the unclipped values.
// One deliberate difference: the argument of acos is clamped to [-1, 1].  For a double root that is not exactly
// diagonal the reference's ratio R/sqrt(-Q^3) can round past 1, acos gives NaN and the element silently reports
// 0 strain (:50-58,61-74); here it reports the eigenvalues.  Where the reference is finite the two agree.
FTB_HD void principal_strains(const double csum[6], double* smax, double* smin, double* shear, const int countGP = 8) {
  const double pre = 0.5 / (double)countGP;
  const double a = pre * csum[0] - 0.5, d = pre * csum[1] - 0.5, f = pre * csum[2] - 0.5;
  const double e = pre * csum[3], c = pre * csum[4], b = pre * csum[5];
  const double p1 = b * b + c * c + e * e;
  double eps1, eps2, eps3;
  if (p1 == 0) {
    eps1 = a; eps2 = d; eps3 = f;
  } else {
    double I1 = a + d + f;
    const double I2 = a * (d + f) + d * f - b * b - c * c - e * e;
    const double I3 = a * d * f + 2.0 * b * c * e - b * b * f - c * c * d - e * e * a;
    const double Q = (3.0 * I2 - I1 * I1) / 9.0;
    const double R = (2.0 * I1 * I1 * I1 - 9.0 * I1 * I2 + 27.0 * I3) / 54.0;
    double ratio = R / sqrt(-Q * Q * Q);
    ratio = ratio > 1.0 ? 1.0 : (ratio < -1.0 ? -1.0 : ratio);
    const double theta = acos(ratio);
    const double sqrtQ = 2.0 * sqrt(-Q);
    const double kPi = 3.14159265358979323846;  // 4 atan(1)
    I1 = I1 / 3.0;
    eps1 = sqrtQ * cos(theta / 3.0) + I1;
    eps2 = sqrtQ * cos((theta + 2.0 * kPi) / 3.0) + I1;
    eps3 = sqrtQ * cos((theta + 4.0 * kPi) / 3.0) + I1;
  }
  const double mx = fmax(eps3, fmax(eps2, eps1)), mn = fmin(eps3, fmin(eps2, eps1));
  *shear = 0.5 * (mx - mn);
  *smax = mx > 0.0 ? mx : 0.0;
  *smin = mn < 0.0 ? mn : 0.0;
}

// volumeHexahedron (Geometry.cpp:3-27) from the mode vectors of the nodal coordinates
FTB_HD double hex_volume_modes(const double xm[7][3]) {
  const double* g1 = xm[0]; const double* g2 = xm[1]; const double* g3 = xm[2];
  const double* g12 = xm[3]; const double* g23 = xm[4]; const double* g13 = xm[5];
  return (tp3(g12, g2, g23) + tp3(g1, g12, g13) + tp3(g13, g23, g3)) / 192.0 + tp3(g1, g2, g3) / 64.0;
}

// The whole element: nodal X[8][3], U[8][3] (C3D8 node order) -> fe[8][3].
// MATSEL: compile-time material id (1..5) when the whole launch is uniform,
// or -1 for the generic per-element switch.  Returns status bits:
// 1 = unknown material, 2 = non-positive det J0, 4 = non-finite / non-positive det F.
// Nodal input of an element: get(c, nx, nu) delivers component c of the 8 reference coordinates and displacements.
// ArrayIn reads plain arrays; the force kernel passes a staged source (k_elem) whose later components arrive
// through shared memory.
struct ArrayIn {
  const double (*X)[3];
  const double (*U)[3];
  FTB_HD void get(const int c, double nx[8], double nu[8]) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) { nx[k] = X[k][c]; nu[k] = U[k][c]; }
  }
};
// scratch slot in which node k of component c is parked before that component's columns are built (the 24 column
// slots of component c are free until then)
#define FTB_STAGE_SLOT(f, k, c) FTB_COL(f, (k) >> 2, (k) & 3, c)

template <int MATSEL, bool WITH_DT, class In, class Hist, class Out, class Scratch>
FTB_HD int hex8_element_in(const In& in, int mat, const double* __restrict__ mp, const bool updHist, const Hist& hist,
                           const Out& out, Scratch& S, double fe[8][3], double* dtElem);

template <int MATSEL, bool WITH_DT, class Hist, class Out, class Scratch>
FTB_HD int hex8_element(const double X[8][3], const double U[8][3], int mat, const double* __restrict__ mp,
                        const bool updHist, const Hist& hist, const Out& out, Scratch& S, double fe[8][3],
                        double* dtElem) {
  return hex8_element_in<MATSEL, WITH_DT>(ArrayIn{X, U}, mat, mp, updHist, hist, out, S, fe, dtElem);
}

template <int MATSEL, bool WITH_DT, class In, class Hist, class Out, class Scratch>
FTB_HD int hex8_element_in(const In& in, int mat, const double* __restrict__ mp, const bool updHist, const Hist& hist,
                           const Out& out, Scratch& S, double fe[8][3], double* dtElem) {
  if (MATSEL >= 0) mat = MATSEL;
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
  double dtk = 0.0;  // 1 / (512 A_max c_e)
  {
    // One space component at a time keeps the live set small: 8+8 nodal values -> 7+7 modes ->
    // 24 column entries written to the scratch; only the 21 current-configuration modes (for the
    // time step) survive the loop.
    double xm[7][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double nx[8], nu[8], gX[7], gU[7];
      in.get(c, nx, nu);
      hex_modes(nx, gX);
      hex_modes(nu, gU);
      if (WITH_DT) {
#pragma unroll
        for (int m = 0; m < 7; ++m) xm[m][c] = gX[m] + gU[m];
      }
      // Jacobian columns 8 dX/dxi_t and 8 dU/dxi_t at the Gauss points.  Column xi depends only on the
      // signs (s2,s3) of the point, eta on (s1,s3), zeta on (s1,s2): 12 distinct vectors per field
      // instead of 24, each built by a 4-point butterfly from the modes (bilinear modes x a, trilinear x a^2).
      const double X12 = a * gX[3], X23 = a * gX[4], X13 = a * gX[5], X123 = a2 * gX[6];
      col_butterfly(S, 0, 0, c, gX[0], X12, X13, X123);  // xi  : (s2, s3)
      col_butterfly(S, 0, 1, c, gX[1], X12, X23, X123);  // eta : (s1, s3)
      col_butterfly(S, 0, 2, c, gX[2], X13, X23, X123);  // zeta: (s1, s2)
      const double U12 = a * gU[3], U23 = a * gU[4], U13 = a * gU[5], U123 = a2 * gU[6];
      col_butterfly(S, 1, 0, c, gU[0], U12, U13, U123);
      col_butterfly(S, 1, 1, c, gU[1], U12, U23, U123);
      col_butterfly(S, 1, 2, c, gU[2], U13, U23, U123);
    }
    // CalculateTimeStep.cpp:19: dt = (V / A_max) / c_e.  The current volume is not taken from volumeHexahedron's four
    // triple products (Geometry.cpp:3-27) but from the Gauss loop below: V = sum_gp det F det J0, exact for the trilinear
    // hexahedron (2 x 2 x 2 points integrate the triquadratic Jacobian exactly) and free: det F is needed anyway.
    if (WITH_DT) dtk = 1.0 / (512.0 * hex_face_amax(xm) * mp[MP_CE]);
  }
  int status = 0;
  double vsum = 0.0;
  double phi[7][3];
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) phi[m][c] = 0.0;

  // The Gauss-point loop is deliberately NOT unrolled: one iteration has 9-wide instruction-level
  // parallelism (3x3 blocks), the rolled body fits the instruction cache, and the register count
  // stays low enough for 12-16 resident warps per SM.  Signs are run-time +-1.0 folded into FMAs.
#ifndef FTB_GP_UNROLL
#define FTB_GP_UNROLL 1
#endif
  constexpr int kGpUnroll = FTB_GP_UNROLL;
#if defined(__CUDA_ARCH__)
#pragma unroll kGpUnroll
#endif
  for (int gp = 0; gp < 8; ++gp) {
    // reference numbering (GaussQuadrature3D.cpp:19-49): xi + for gp 1,2,5,6; eta + for 2,3,6,7; zeta + for 0..3
    const int b1 = ((gp + 1) >> 1) & 1, b2 = (gp >> 1) & 1, b3 = ((gp >> 2) & 1) ^ 1;
    const double s1 = b1 ? 1.0 : -1.0, s2 = b2 ? 1.0 : -1.0, s3 = b3 ? 1.0 : -1.0;
    const double s23 = s2 * s3, s13 = s1 * s3, s12 = s1 * s2;
    const int qx = b2 + 2 * b3, qe = b1 + 2 * b3, qz = b1 + 2 * b2;
    double J0[3][3], Uh[3][3], cJ[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      J0[i][0] = S.ld(FTB_COL(0, 0, qx, i));
      J0[i][1] = S.ld(FTB_COL(0, 1, qe, i));
      J0[i][2] = S.ld(FTB_COL(0, 2, qz, i));
      Uh[i][0] = S.ld(FTB_COL(1, 0, qx, i));
      Uh[i][1] = S.ld(FTB_COL(1, 1, qe, i));
      Uh[i][2] = S.ld(FTB_COL(1, 2, qz, i));
    }
    cofactor3(J0, cJ);
    const double det = J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2];  // 512 detJ0
    if (!(det > 0.0)) status |= 2;
    const double rdet = ftb_rcp(det);
    // F = I + Uh * J0^-1 ,  J0^-1[c][j] = cJ[j][c] / det
    double F[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        F[i][j] = (Uh[i][0] * cJ[j][0] + Uh[i][1] * cJ[j][1] + Uh[i][2] * cJ[j][2]) * rdet + (i == j ? 1.0 : 0.0);
    double cF[3][3];
    cofactor3(F, cF);
    const double J = F[0][0] * cF[0][0] + F[0][1] * cF[0][1] + F[0][2] * cF[0][2];
    if (!(J > 0.0) && mat != 0) status |= 4;
    if (WITH_DT) vsum = fma(J, det, vsum);  // 512 x the current volume of this point's octant
    double P[3][3], Sv[6];
    GpHistory h;
    if (mat == 5) hist.load(gp, h);
    status |= material_P<Out::want_S>(mat, F, cF, J, mp, &h, updHist, P, Sv);
    if (mat == 5 && updHist) hist.store(gp, h);
    if (Out::enabled) out.put(gp, F, J, Sv);
    // Q = P * cof(J0)  (unscaled: 64 x the true one), accumulated into the 7 force modes
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double Q0 = P[i][0] * cJ[0][0] + P[i][1] * cJ[1][0] + P[i][2] * cJ[2][0];
      const double Q1 = P[i][0] * cJ[0][1] + P[i][1] * cJ[1][1] + P[i][2] * cJ[2][1];
      const double Q2 = P[i][0] * cJ[0][2] + P[i][1] * cJ[1][2] + P[i][2] * cJ[2][2];
      phi[0][i] += Q0;
      phi[1][i] += Q1;
      phi[2][i] += Q2;
      phi[3][i] = fma(s2, Q0, fma(s1, Q1, phi[3][i]));
      phi[4][i] = fma(s3, Q1, fma(s2, Q2, phi[4][i]));
      phi[5][i] = fma(s3, Q0, fma(s1, Q2, phi[5][i]));
      phi[6][i] = fma(s23, Q0, fma(s13, Q1, fma(s12, Q2, phi[6][i])));
    }
  }
  if (WITH_DT) *dtElem = vsum * dtk;
  // scale: 1/8 (dN/dxi) * 1/64 (cofactor of 8 J0); bilinear modes carry a, trilinear a^2
  const double w0 = 1.0 / 512.0, w1 = a / 512.0, w2 = a2 / 512.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double p[7], f[8];
    p[0] = phi[0][c] * w0; p[1] = phi[1][c] * w0; p[2] = phi[2][c] * w0;
    p[3] = phi[3][c] * w1; p[4] = phi[4][c] * w1; p[5] = phi[5][c] * w1;
    p[6] = phi[6][c] * w2;
    hex_modes_to_nodes(p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) fe[k][c] = f[k];
  }
  return status;
}

// ---------------------------------------------------------------------------------------------
// Hexahedra with an AFFINE reference geometry (parallelepipeds: the four edges along each local direction are the
// same vector, bit for bit -- every element of a structured / voxel mesh).  dX/dxi is then the same matrix at all
// eight Gauss points, so cof(J0), det J0 and J0^-1 are formed once per element instead of once per point, the
// reference coordinates need no Jacobian columns (and only nodes 0, 1, 3, 4 are read), and
//   F = I + (dU/dxi) J0^-1      costs 27 fused multiply-adds per point instead of 36 + cofactors + a division.
// Same formulas otherwise; for such an element the general path's Jacobian columns reduce to exactly these values
// (the bilinear and trilinear coordinate modes are exact zeros), so the two paths differ by rounding only
// (tests/test_element_math_cpu.py pins 1e-13).  The host decides per run of elements which kernel is launched.
//
// Scratch layout (54 doubles): 36 dU/dxi column entries, then cof(8 J0) (9) and (8 J0)^-1 (9).
#define FTB_ACOL(t, q, c) (((t) * 4 + (q)) * 3 + (c))
#define FTB_ACJ(j, c) (36 + (j) * 3 + (c))
#define FTB_AJI(c, j) (45 + (c) * 3 + (j))
// staging slots of the kernel's asynchronous gather: node k of displacement component c waits in a column slot of
// the same component; reference node kk (0..3 = C3D8 nodes 0, 1, 3, 4) of component c = 1, 2 in the cofactor slots
#define FTB_ASTAGE_U(k, c) FTB_ACOL((k) >> 2, (k) & 3, c)
#define FTB_ASTAGE_X(kk, c) (36 + ((c) - 1) * 4 + (kk))
#define FTB_AFFINE_SLOTS 54
#ifndef FTB_AFF_GP_UNROLL
#define FTB_AFF_GP_UNROLL 1
#endif

struct LocalScratchAffine {
  double v[FTB_AFFINE_SLOTS];
  FTB_HD void st(int i, double x) { v[i] = x; }
  FTB_HD double ld(int i) const { return v[i]; }
  FTB_HD double ld_inloop(int i) const { return v[i]; }  // the kernel's scratch re-reads loop invariants from shared memory
};
// nodal input of the affine path: getX delivers component c of nodes 0, 1, 3, 4; getU of all eight displacements
struct ArrayInAffine {
  const double (*X)[3];
  const double (*U)[3];
  FTB_HD void getX(const int c, double x[4]) const { x[0] = X[0][c]; x[1] = X[1][c]; x[2] = X[3][c]; x[3] = X[4][c]; }
  FTB_HD void getU(const int c, double nu[8]) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) nu[k] = U[k][c];
  }
};
// exact test the host applies per element (X in C3D8 node order): parallel edges are identical vectors
FTB_HD bool hex8_is_affine(const double X[8][3]) {
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double d1 = X[1][c] - X[0][c], d2 = X[3][c] - X[0][c], d3 = X[4][c] - X[0][c];
    ok = ok && (X[2][c] - X[3][c] == d1) && (X[5][c] - X[4][c] == d1) && (X[6][c] - X[7][c] == d1);
    ok = ok && (X[2][c] - X[1][c] == d2) && (X[7][c] - X[4][c] == d2) && (X[6][c] - X[5][c] == d2);
    ok = ok && (X[5][c] - X[1][c] == d3) && (X[6][c] - X[2][c] == d3) && (X[7][c] - X[3][c] == d3);
  }
  return ok;
}

template <int MATSEL, bool WITH_DT, class In, class Hist, class Out, class Scratch>
FTB_HD int hex8_element_affine_in(const In& in, int mat, const double* __restrict__ mp, const bool updHist,
                                  const Hist& hist, const Out& out, Scratch& S, double fe[8][3], double* dtElem) {
  if (MATSEL >= 0) mat = MATSEL;
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
  int status = 0;
  double dtk = 0.0;
  {
    double xm[7][3];
    double J0[3][3];  // 8 dX/dxi = 4 x edge vectors (exact scaling)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double x[4], nu[8], gU[7];
      in.getX(c, x);
      J0[c][0] = 4.0 * (x[1] - x[0]);
      J0[c][1] = 4.0 * (x[2] - x[0]);
      J0[c][2] = 4.0 * (x[3] - x[0]);
      in.getU(c, nu);
      hex_modes(nu, gU);
      if (WITH_DT) {
        xm[0][c] = J0[c][0] + gU[0]; xm[1][c] = J0[c][1] + gU[1]; xm[2][c] = J0[c][2] + gU[2];
        xm[3][c] = gU[3]; xm[4][c] = gU[4]; xm[5][c] = gU[5]; xm[6][c] = gU[6];
      }
      const double U12 = a * gU[3], U23 = a * gU[4], U13 = a * gU[5], U123 = a2 * gU[6];
      {  // the three displacement-gradient columns at the four sign pairs each (see col_butterfly)
        const double A[3] = {gU[0], gU[1], gU[2]}, B[3] = {U12, U12, U13}, C[3] = {U13, U23, U23};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const double ad = A[t] + U123, am = A[t] - U123, bc = B[t] + C[t], bm = B[t] - C[t];
          S.st(FTB_ACOL(t, 3, c), ad + bc);
          S.st(FTB_ACOL(t, 0, c), ad - bc);
          S.st(FTB_ACOL(t, 1, c), am + bm);
          S.st(FTB_ACOL(t, 2, c), am - bm);
        }
      }
    }
    double cJ[3][3];
    cofactor3(J0, cJ);
    const double det = J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2];  // 512 detJ0
    if (!(det > 0.0)) status |= 2;
    const double rdet = ftb_rcp(det);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        S.st(FTB_ACJ(j, c), cJ[j][c]);
        S.st(FTB_AJI(c, j), cJ[j][c] * rdet);  // J0^-1[c][j] = cof[j][c] / det
      }
    // dt = (V / A_max) / c_e with V = det J0 sum_gp det F (see hex8_element_in); det here is 512 det J0
    if (WITH_DT) dtk = det * ftb_rcp(512.0 * hex_face_amax(xm) * mp[MP_CE]);
  }
  double vsum = 0.0;
  double phi[7][3];
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) phi[m][c] = 0.0;
  // The loop stays rolled (FTB_AFF_GP_UNROLL = 1).  Measured at 100^3, neo-Hookean: with the library log / division rolled
  // 164.0 us, unrolled by 2 161.9, by 4 159.8, by 8 167.8; with ftb_ln_rcp rolled **155.6**, by 2 176, by 4 180 (spills).
  // HGO, HGO + Prony and the injury variant lose with any unrolling.
  constexpr int kGpUnroll = (MATSEL == 1 && !Out::enabled) ? FTB_AFF_GP_UNROLL : 1;
#if defined(__CUDA_ARCH__)
#pragma unroll kGpUnroll
#endif
  for (int gp = 0; gp < 8; ++gp) {
    const int b1 = ((gp + 1) >> 1) & 1, b2 = (gp >> 1) & 1, b3 = ((gp >> 2) & 1) ^ 1;
    const double s1 = b1 ? 1.0 : -1.0, s2 = b2 ? 1.0 : -1.0, s3 = b3 ? 1.0 : -1.0;
    const double s23 = s2 * s3, s13 = s1 * s3, s12 = s1 * s2;
    const int qx = b2 + 2 * b3, qe = b1 + 2 * b3, qz = b1 + 2 * b2;
    double F[3][3];
    {
      double Ji[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) Ji[c][j] = S.ld_inloop(FTB_AJI(c, j));
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double u0 = S.ld(FTB_ACOL(0, qx, i)), u1 = S.ld(FTB_ACOL(1, qe, i)), u2 = S.ld(FTB_ACOL(2, qz, i));
#pragma unroll
        for (int j = 0; j < 3; ++j) F[i][j] = fma(u0, Ji[0][j], fma(u1, Ji[1][j], fma(u2, Ji[2][j], (i == j ? 1.0 : 0.0))));
      }
    }
    double cF[3][3];
    cofactor3(F, cF);
    const double J = F[0][0] * cF[0][0] + F[0][1] * cF[0][1] + F[0][2] * cF[0][2];
    if (!(J > 0.0) && mat != 0) status |= 4;
    if (WITH_DT) vsum += J;
    double P[3][3], Sv[6];
    GpHistory h;
    if (mat == 5) hist.load(gp, h);
    status |= material_P<Out::want_S>(mat, F, cF, J, mp, &h, updHist, P, Sv);
    if (mat == 5 && updHist) hist.store(gp, h);
    if (Out::enabled) out.put(gp, F, J, Sv);
    double cJ[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) cJ[j][c] = S.ld_inloop(FTB_ACJ(j, c));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double Q0 = P[i][0] * cJ[0][0] + P[i][1] * cJ[1][0] + P[i][2] * cJ[2][0];
      const double Q1 = P[i][0] * cJ[0][1] + P[i][1] * cJ[1][1] + P[i][2] * cJ[2][1];
      const double Q2 = P[i][0] * cJ[0][2] + P[i][1] * cJ[1][2] + P[i][2] * cJ[2][2];
      phi[0][i] += Q0;
      phi[1][i] += Q1;
      phi[2][i] += Q2;
      phi[3][i] = fma(s2, Q0, fma(s1, Q1, phi[3][i]));
      phi[4][i] = fma(s3, Q1, fma(s2, Q2, phi[4][i]));
      phi[5][i] = fma(s3, Q0, fma(s1, Q2, phi[5][i]));
      phi[6][i] = fma(s23, Q0, fma(s13, Q1, fma(s12, Q2, phi[6][i])));
    }
  }
  if (WITH_DT) *dtElem = vsum * dtk;
  const double w0 = 1.0 / 512.0, w1 = a / 512.0, w2 = a2 / 512.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double p[7], f[8];
    p[0] = phi[0][c] * w0; p[1] = phi[1][c] * w0; p[2] = phi[2][c] * w0;
    p[3] = phi[3][c] * w1; p[4] = phi[4][c] * w1; p[5] = phi[5][c] * w1;
    p[6] = phi[6][c] * w2;
    hex_modes_to_nodes(p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) fe[k][c] = f[k];
  }
  return status;
}

// ---------------------------------------------------------------------------------------------
// The neo-Hookean parallelepiped in CURRENT-JACOBIAN form (round 2; the default for runs of material-1 parallelepipeds
// without the strain outputs).  Same element, same Gauss rule, algebra rearranged so that the Gauss loop only carries
// what is nonlinear:
//   Ft = dx/dxi = J0 + dU/dxi (the Jacobian of the CURRENT configuration; F = Ft J0^-1 is never formed),
//   cof F = cof(Ft) cof(J0)^-1,  J = det Ft / det J0,
//   Q = P cof(J0) = mu Ft M + c cof(Ft),   M = J0^-1 cof(J0) = cof(J0)^T cof(J0) / det J0 (symmetric, one per element),
//   c = (lambda ln J - mu) / J   (CompressibleNeoHookean.cpp:43-55 pushed through P = F S).
// The first term is LINEAR in the nodal positions and M is the same at all eight points, so its sum over the Gauss
// points is taken in closed form in the mode basis (sign patterns of different modes are orthogonal over the 2x2x2
// rule): linear modes 8 mu xm_t M, bilinear 8 a mu (...), trilinear 8 a^2 mu xm_123 tr M -- about 60 flops per element
// instead of 63 per point.  What stays in the loop is cof(Ft), det, ln J, 1/J and 36 accumulations of c cof(Ft): 9 shared
// loads and ~95 fp64 instructions per point against 27 loads and ~172 for hex8_element_affine_in, no J0^-1 / cof(J0)
// slots.  The rest state is no longer an exact zero (8 mu J0 M and sum c cof(Ft) cancel to rounding, a spurious
// strain of ~1e-16 -- the reference, which builds F from coordinates, has the same noise); differs from
// hex8_element_affine_in by rounding only (tests/test_element_math_cpu.py pins 1e-12 of the force maximum at 0.4 %
// strain).
//
// ln J with a short series first: |s| <= 1/16 (0.882 <= J <= 1.133) needs 6 terms for 1e-18.
FTB_HD void ftb_ln_rcp_tiered(const double J, double* lnJ, double* rJ) {
#if defined(__CUDA_ARCH__) && !defined(FTB_LIBM_MATERIAL)
  const double w = J - 1.0, t = J + 1.0;
  const double rt = ftb_rcp(t);
  double s = w * rt;
  s = fma(rt, fma(-s, t, w), s);
  *rJ = ftb_rcp(J);
  const double z = s * s, zz = z * z, s2 = s + s;
  if (fabs(s) <= 0.0625) {
    // 1/3 + z/5 + z^2/7 + z^3/9 + z^4/11 + z^5/13: even and odd powers as two chains
    const double pe = fma(fma(FTB_LN_C[4], zz, FTB_LN_C[2]), zz, FTB_LN_C[0]);
    const double po = fma(fma(FTB_LN_C[5], zz, FTB_LN_C[3]), zz, FTB_LN_C[1]);
    *lnJ = fma(s2 * z, fma(po, z, pe), s2);
  } else if (fabs(s) <= 0.25) {
    double pe = FTB_LN_C[12], po = FTB_LN_C[11];
#pragma unroll
    for (int k = 10; k >= 0; k -= 2) pe = fma(pe, zz, FTB_LN_C[k]);
#pragma unroll
    for (int k = 9; k >= 1; k -= 2) po = fma(po, zz, FTB_LN_C[k]);
    *lnJ = fma(s2 * z, fma(po, z, pe), s2);
  } else {
    *lnJ = log(J);
  }
#else
  *lnJ = log(J);
  *rJ = 1.0 / J;
#endif
}
#define FTB_NH_SLOTS 44  // 36 column entries + the staging slots of the reference nodes (FTB_ASTAGE_X)
#ifndef FTB_NH_GP_UNROLL
#define FTB_NH_GP_UNROLL 8  // measured at 100^3: rolled 121.0 us, by 2 116.6, by 4 113.6, by 8 112.5 (profiles/r02_k_elem_affine_nh_variants.txt)
#endif
#ifndef FTB_CJ4_GP_UNROLL
#define FTB_CJ4_GP_UNROLL 1
#endif
// MAT = 1: as above.  MAT = 4 (HGO with isotropic fibre dispersion, HGOIsotropic.cpp:44-84): the Cauchy stress is
// sigma = pref dev(B) + hydro I with B = F F^T, and B cof F = F (F^T cof F) = J F, so
//   P = sigma cof F = (pref J) F + (hydro - pref tr B / 3) cof F,    Q = alpha Ft M + gamma cof(Ft)
// -- the neo-Hookean structure with coefficients that vary from point to point (alpha through J and I1), hence no
// closed-form sum: G = Ft M is formed per point (M in 6 scratch slots) and also delivers tr B = (G : Ft) / det J0.
// 63 multiply-adds per point for F, B, P and Q become 45, and the 18 loads of cof(J0), J0^-1 become 6.
#define FTB_CJ_M(k) (36 + (k))  // M00 M11 M22 M12 M02 M01 (Voigt) in the staging slots of the reference nodes, free by then
template <int MAT, bool WITH_DT, class In, class Scratch>
FTB_HD int hex8_element_affine_cj(const In& in, const double* __restrict__ mp, Scratch& S, double fe[8][3], double* dtElem) {
  static_assert(MAT == 1 || MAT == 4, "current-Jacobian form: neo-Hookean and HGO");
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
  const double mu = mp[MP_MU], lambda = mp[MP_LAMBDA];
  int status = 0;
  double dtk = 0.0, rdet0;
  double phi[7][3];
  {
    double xm[7][3];  // modes of the current coordinates (8 dx/dxi scaling): linear ones J0 + gU, the rest gU
    double J0[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double x[4], nu[8], gU[7];
      in.getX(c, x);
      J0[c][0] = 4.0 * (x[1] - x[0]);
      J0[c][1] = 4.0 * (x[2] - x[0]);
      J0[c][2] = 4.0 * (x[3] - x[0]);
      in.getU(c, nu);
      hex_modes(nu, gU);
      xm[0][c] = J0[c][0] + gU[0]; xm[1][c] = J0[c][1] + gU[1]; xm[2][c] = J0[c][2] + gU[2];
      xm[3][c] = gU[3]; xm[4][c] = gU[4]; xm[5][c] = gU[5]; xm[6][c] = gU[6];
      const double U12 = a * gU[3], U23 = a * gU[4], U13 = a * gU[5], U123 = a2 * gU[6];
      const double A[3] = {xm[0][c], xm[1][c], xm[2][c]}, B[3] = {U12, U12, U13}, C[3] = {U13, U23, U23};
#pragma unroll
      for (int t = 0; t < 3; ++t) {  // column t of 8 dx/dxi at its four sign pairs (layout of hex8_element_affine_in)
        const double ad = A[t] + U123, am = A[t] - U123, bc = B[t] + C[t], bm = B[t] - C[t];
        S.st(FTB_ACOL(t, 3, c), ad + bc);
        S.st(FTB_ACOL(t, 0, c), ad - bc);
        S.st(FTB_ACOL(t, 1, c), am + bm);
        S.st(FTB_ACOL(t, 2, c), am - bm);
      }
    }
    double cJ[3][3];
    cofactor3(J0, cJ);
    const double det = J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2];  // 512 det J0
    if (!(det > 0.0)) status |= 2;
    rdet0 = ftb_rcp(det);
    // dt = (V / A_max) / c_e with V = det J0 sum_gp det F = sum_gp det Ft (see hex8_element_in); before the linear term,
    // which turns the modes into the force accumulators component by component
    if (WITH_DT) dtk = det * ftb_rcp(512.0 * hex_face_amax(xm) * mp[MP_CE]);
    // MAT 1: 8 mu M;  MAT 4: M itself (M = cof^T cof / det)
    const double k8 = (MAT == 1 ? 8.0 * mu : 1.0) * rdet0;
    const double M00 = k8 * (cJ[0][0] * cJ[0][0] + cJ[1][0] * cJ[1][0] + cJ[2][0] * cJ[2][0]);
    const double M11 = k8 * (cJ[0][1] * cJ[0][1] + cJ[1][1] * cJ[1][1] + cJ[2][1] * cJ[2][1]);
    const double M22 = k8 * (cJ[0][2] * cJ[0][2] + cJ[1][2] * cJ[1][2] + cJ[2][2] * cJ[2][2]);
    const double M01 = k8 * (cJ[0][0] * cJ[0][1] + cJ[1][0] * cJ[1][1] + cJ[2][0] * cJ[2][1]);
    const double M02 = k8 * (cJ[0][0] * cJ[0][2] + cJ[1][0] * cJ[1][2] + cJ[2][0] * cJ[2][2]);
    const double M12 = k8 * (cJ[0][1] * cJ[0][2] + cJ[1][1] * cJ[1][2] + cJ[2][1] * cJ[2][2]);
    if (MAT == 1) {
      const double aM3 = a * (M00 + M11), aM4 = a * (M11 + M22), aM5 = a * (M00 + M22), aM6 = a2 * (M00 + M11 + M22);
      const double aM01 = a * M01, aM02 = a * M02, aM12 = a * M12;
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // sum over the Gauss points of mu Ft M against the mode gradients, in closed form
        phi[0][i] = xm[0][i] * M00 + xm[1][i] * M01 + xm[2][i] * M02;
        phi[1][i] = xm[0][i] * M01 + xm[1][i] * M11 + xm[2][i] * M12;
        phi[2][i] = xm[0][i] * M02 + xm[1][i] * M12 + xm[2][i] * M22;
        phi[3][i] = xm[3][i] * aM3 + xm[4][i] * aM02 + xm[5][i] * aM12;
        phi[4][i] = xm[4][i] * aM4 + xm[5][i] * aM01 + xm[3][i] * aM02;
        phi[5][i] = xm[5][i] * aM5 + xm[4][i] * aM01 + xm[3][i] * aM12;
        phi[6][i] = xm[6][i] * aM6;
      }
    } else {
      S.st(FTB_CJ_M(0), M00); S.st(FTB_CJ_M(1), M11); S.st(FTB_CJ_M(2), M22);
      S.st(FTB_CJ_M(3), M12); S.st(FTB_CJ_M(4), M02); S.st(FTB_CJ_M(5), M01);
#pragma unroll
      for (int m = 0; m < 7; ++m)
#pragma unroll
        for (int i = 0; i < 3; ++i) phi[m][i] = 0.0;
    }
  }
  double vsum = 0.0;
  constexpr int kGpUnroll = MAT == 1 ? FTB_NH_GP_UNROLL : FTB_CJ4_GP_UNROLL;
#if defined(__CUDA_ARCH__)
#pragma unroll kGpUnroll
#endif
  for (int gp = 0; gp < 8; ++gp) {
    const int b1 = ((gp + 1) >> 1) & 1, b2 = (gp >> 1) & 1, b3 = ((gp >> 2) & 1) ^ 1;
    const int qx = b2 + 2 * b3, qe = b1 + 2 * b3, qz = b1 + 2 * b2;
    double Ft[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // ld_inloop: a load the compiler cannot replace by the value stored above (it would keep all 36 in registers)
      Ft[i][0] = S.ld_inloop(FTB_ACOL(0, qx, i));
      Ft[i][1] = S.ld_inloop(FTB_ACOL(1, qe, i));
      Ft[i][2] = S.ld_inloop(FTB_ACOL(2, qz, i));
    }
    double cF[3][3];
    cofactor3(Ft, cF);
    const double det = Ft[0][0] * cF[0][0] + Ft[0][1] * cF[0][1] + Ft[0][2] * cF[0][2];
    if (!(det > 0.0)) status |= 4;
    const double J = det * rdet0;
    if (WITH_DT) vsum += J;
    if (MAT == 1) {
      double lnJ, rJ;
      ftb_ln_rcp_tiered(J, &lnJ, &rJ);
      const double cc = (lambda * lnJ - mu) * rJ;
      const double c1 = b1 ? cc : -cc, c2 = b2 ? cc : -cc, c3 = b3 ? cc : -cc;
      const double c23 = (b2 == b3) ? cc : -cc, c13 = (b1 == b3) ? cc : -cc, c12 = (b1 == b2) ? cc : -cc;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        phi[0][i] = fma(cc, cF[i][0], phi[0][i]);
        phi[1][i] = fma(cc, cF[i][1], phi[1][i]);
        phi[2][i] = fma(cc, cF[i][2], phi[2][i]);
        phi[3][i] = fma(c2, cF[i][0], fma(c1, cF[i][1], phi[3][i]));
        phi[4][i] = fma(c3, cF[i][1], fma(c2, cF[i][2], phi[4][i]));
        phi[5][i] = fma(c3, cF[i][0], fma(c1, cF[i][2], phi[5][i]));
        phi[6][i] = fma(c23, cF[i][0], fma(c13, cF[i][1], fma(c12, cF[i][2], phi[6][i])));
      }
    } else {
      const double M00 = S.ld_inloop(FTB_CJ_M(0)), M11 = S.ld_inloop(FTB_CJ_M(1)), M22 = S.ld_inloop(FTB_CJ_M(2));
      const double M12 = S.ld_inloop(FTB_CJ_M(3)), M02 = S.ld_inloop(FTB_CJ_M(4)), M01 = S.ld_inloop(FTB_CJ_M(5));
      double G[3][3];
      double trB = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        G[i][0] = Ft[i][0] * M00 + Ft[i][1] * M01 + Ft[i][2] * M02;
        G[i][1] = Ft[i][0] * M01 + Ft[i][1] * M11 + Ft[i][2] * M12;
        G[i][2] = Ft[i][0] * M02 + Ft[i][1] * M12 + Ft[i][2] * M22;
        trB += G[i][0] * Ft[i][0] + G[i][1] * Ft[i][1] + G[i][2] * Ft[i][2];
      }
      trB *= rdet0;
      // the scalars of hgo_cauchy (HGOIsotropic.cpp:44-84), same expressions
      const double k1 = mp[MP_K1], k2 = mp[MP_K2], K = mp[MP_KBULK];
      const double rJ = ftb_rcp(J);
      const double hydro = 0.5 * K * (J * J - 1.0) * rJ;
      const double rc = ftb_rcbrt(J);
      const double Jm23 = rc * rc;
      const double I1 = Jm23 * trB;
      const double kappa = 1.0 / 3.0;
      const double Ea = kappa * (I1 - 3.0);
      double fiber = 0.0;
      if (Ea > 0.0) {
        const double ex = (k2 == 0.0) ? 1.0 : exp(k2 * Ea * Ea);
        fiber = 2.0 * k1 * ex * Ea * kappa;
      }
      const double pref = Jm23 * (mu + fiber) * rJ;
      const double alpha = pref * J, gamma = hydro - pref * (trB * (1.0 / 3.0));
      const double s1 = b1 ? 1.0 : -1.0, s2 = b2 ? 1.0 : -1.0, s3 = b3 ? 1.0 : -1.0;
      const double s23 = s2 * s3, s13 = s1 * s3, s12 = s1 * s2;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double Q0 = fma(alpha, G[i][0], gamma * cF[i][0]);
        const double Q1 = fma(alpha, G[i][1], gamma * cF[i][1]);
        const double Q2 = fma(alpha, G[i][2], gamma * cF[i][2]);
        phi[0][i] += Q0;
        phi[1][i] += Q1;
        phi[2][i] += Q2;
        phi[3][i] = fma(s2, Q0, fma(s1, Q1, phi[3][i]));
        phi[4][i] = fma(s3, Q1, fma(s2, Q2, phi[4][i]));
        phi[5][i] = fma(s3, Q0, fma(s1, Q2, phi[5][i]));
        phi[6][i] = fma(s23, Q0, fma(s13, Q1, fma(s12, Q2, phi[6][i])));
      }
    }
  }
  if (WITH_DT) *dtElem = vsum * dtk;
  const double w0 = 1.0 / 512.0, w1 = a / 512.0, w2 = a2 / 512.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double p[7], f[8];
    p[0] = phi[0][c] * w0; p[1] = phi[1][c] * w0; p[2] = phi[2][c] * w0;
    p[3] = phi[3][c] * w1; p[4] = phi[4][c] * w1; p[5] = phi[5][c] * w1;
    p[6] = phi[6][c] * w2;
    hex_modes_to_nodes(p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) fe[k][c] = f[k];
  }
  return status;
}

template <bool WITH_DT, class In, class Scratch>
FTB_HD int hex8_element_affine_nh(const In& in, const double* __restrict__ mp, Scratch& S, double fe[8][3], double* dtElem) {
  return hex8_element_affine_cj<1, WITH_DT>(in, mp, S, fe, dtElem);
}

// ---------------------------------------------------------------------------------------------
// The parallelepiped element once more, for the brick kernel (k_brick): the same arithmetic as hex8_element_affine_in,
// cut into the two phases of that kernel's software pipeline and with the scratch split by access pattern:
//   hex8_brick_setup  nodal gather -> displacement modes -> the 36 dU/dxi column entries (st_col: the kernel keeps them
//                     in TENSOR MEMORY, 72 columns of the thread's own lane), (8 J0)^-1 (st_ji: 9 shared-memory slots
//                     that are re-read inside the loop), and the loop-invariant part of the element's stable dt;
//   hex8_brick_loop   the eight Gauss points and the inverse butterfly.
// cof(8 J0)[j][c] = det (8 J0)^-1[c][j], so only the inverse is kept and the determinant is folded into the quadrature
// weights of the final butterfly.  Differs from hex8_element_affine_in by rounding only (tests/test_element_math_cpu.py).
struct LocalScratchBrick {
  double col[36], ji[9];
  FTB_HD void st_col(int i, double x) { col[i] = x; }
  FTB_HD void cols_written() const {}
  FTB_HD void ld_cols9(const int idx[9], double out[9]) const {  // the nine column entries of one Gauss point
    for (int k = 0; k < 9; ++k) out[k] = col[idx[k]];
  }
  FTB_HD void st_ji(int i, double x) { ji[i] = x; }
  FTB_HD double ld_ji(int i) const { return ji[i]; }
};

template <class In, class Scratch>
FTB_HD int hex8_brick_setup(const In& in, const double* __restrict__ mp, Scratch& S, double* det_out, double* dtk_out) {
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
  int status = 0;
  double xm[7][3];
  double J0[3][3];  // 8 dX/dxi = 4 x edge vectors (exact scaling)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double x[4], nu[8], gU[7];
    in.getX(c, x);
    J0[c][0] = 4.0 * (x[1] - x[0]);
    J0[c][1] = 4.0 * (x[2] - x[0]);
    J0[c][2] = 4.0 * (x[3] - x[0]);
    in.getU(c, nu);
    hex_modes(nu, gU);
    xm[0][c] = J0[c][0] + gU[0]; xm[1][c] = J0[c][1] + gU[1]; xm[2][c] = J0[c][2] + gU[2];
    xm[3][c] = gU[3]; xm[4][c] = gU[4]; xm[5][c] = gU[5]; xm[6][c] = gU[6];
    const double U12 = a * gU[3], U23 = a * gU[4], U13 = a * gU[5], U123 = a2 * gU[6];
    {
      const double A[3] = {gU[0], gU[1], gU[2]}, B[3] = {U12, U12, U13}, C[3] = {U13, U23, U23};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const double ad = A[t] + U123, am = A[t] - U123, bc = B[t] + C[t], bm = B[t] - C[t];
        S.st_col(FTB_ACOL(t, 3, c), ad + bc);
        S.st_col(FTB_ACOL(t, 0, c), ad - bc);
        S.st_col(FTB_ACOL(t, 1, c), am + bm);
        S.st_col(FTB_ACOL(t, 2, c), am - bm);
      }
    }
  }
  S.cols_written();
  double cJ[3][3];
  cofactor3(J0, cJ);
  const double det = J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2];  // 512 detJ0
  if (!(det > 0.0)) status |= 2;
  const double rdet = ftb_rcp(det);
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) S.st_ji(c * 3 + j, cJ[j][c] * rdet);  // J0^-1[c][j] = cof[j][c] / det
  *det_out = det;
  *dtk_out = det * ftb_rcp(512.0 * hex_face_amax(xm) * mp[MP_CE]);  // dt = (V / A_max) / c_e with V = det J0 sum_gp det F
  return status;
}

template <int MATSEL, class Hist, class Out, class Scratch>
FTB_HD int hex8_brick_loop(int mat, const double* __restrict__ mp, const bool updHist, const Hist& hist, const Out& out,
                           const Scratch& S, const double det, const double dtk, double fe[8][3], double* dtElem) {
  if (MATSEL >= 0) mat = MATSEL;
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
  int status = 0;
  double vsum = 0.0;
  double phi[7][3];
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) phi[m][c] = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int gp = 0; gp < 8; ++gp) {
    const int b1 = ((gp + 1) >> 1) & 1, b2 = (gp >> 1) & 1, b3 = ((gp >> 2) & 1) ^ 1;
    const double s1 = b1 ? 1.0 : -1.0, s2 = b2 ? 1.0 : -1.0, s3 = b3 ? 1.0 : -1.0;
    const double s23 = s2 * s3, s13 = s1 * s3, s12 = s1 * s2;
    const int qx = b2 + 2 * b3, qe = b1 + 2 * b3, qz = b1 + 2 * b2;
    double F[3][3];
    {
      const int idx[9] = {FTB_ACOL(0, qx, 0), FTB_ACOL(1, qe, 0), FTB_ACOL(2, qz, 0), FTB_ACOL(0, qx, 1), FTB_ACOL(1, qe, 1),
                          FTB_ACOL(2, qz, 1), FTB_ACOL(0, qx, 2), FTB_ACOL(1, qe, 2), FTB_ACOL(2, qz, 2)};
      double Ji[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) Ji[c][j] = S.ld_ji(c * 3 + j);
      double uc[9];
      S.ld_cols9(idx, uc);  // (issued behind the shared-memory loads: their latency passes under the tensor-memory round trip)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          F[i][j] = fma(uc[3 * i], Ji[0][j], fma(uc[3 * i + 1], Ji[1][j], fma(uc[3 * i + 2], Ji[2][j], (i == j ? 1.0 : 0.0))));
    }
    double cF[3][3];
    cofactor3(F, cF);
    const double J = F[0][0] * cF[0][0] + F[0][1] * cF[0][1] + F[0][2] * cF[0][2];
    if (!(J > 0.0) && mat != 0) status |= 4;
    vsum += J;
    double P[3][3], Sv[6];
    GpHistory h;
    if (mat == 5) hist.load(gp, h);
    status |= material_P<Out::want_S>(mat, F, cF, J, mp, &h, updHist, P, Sv);
    if (mat == 5 && updHist) hist.store(gp, h);
    if (Out::enabled) out.put(gp, F, J, Sv);
    double Ji[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 3; ++j) Ji[c][j] = S.ld_ji(c * 3 + j);
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // Q / det = P J0^-T
      const double Q0 = P[i][0] * Ji[0][0] + P[i][1] * Ji[0][1] + P[i][2] * Ji[0][2];
      const double Q1 = P[i][0] * Ji[1][0] + P[i][1] * Ji[1][1] + P[i][2] * Ji[1][2];
      const double Q2 = P[i][0] * Ji[2][0] + P[i][1] * Ji[2][1] + P[i][2] * Ji[2][2];
      phi[0][i] += Q0;
      phi[1][i] += Q1;
      phi[2][i] += Q2;
      phi[3][i] = fma(s2, Q0, fma(s1, Q1, phi[3][i]));
      phi[4][i] = fma(s3, Q1, fma(s2, Q2, phi[4][i]));
      phi[5][i] = fma(s3, Q0, fma(s1, Q2, phi[5][i]));
      phi[6][i] = fma(s23, Q0, fma(s13, Q1, fma(s12, Q2, phi[6][i])));
    }
  }
  *dtElem = vsum * dtk;
  const double w0 = det * (1.0 / 512.0), w1 = det * (a / 512.0), w2 = det * (a2 / 512.0);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double p[7], f[8];
    p[0] = phi[0][c] * w0; p[1] = phi[1][c] * w0; p[2] = phi[2][c] * w0;
    p[3] = phi[3][c] * w1; p[4] = phi[4][c] * w1; p[5] = phi[5][c] * w1;
    p[6] = phi[6][c] * w2;
    hex_modes_to_nodes(p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) fe[k][c] = f[k];
  }
  return status;
}

template <int MATSEL, class In, class Hist, class Out, class Scratch>
FTB_HD int hex8_element_brick_in(const In& in, int mat, const double* __restrict__ mp, const bool updHist,
                                 const Hist& hist, const Out& out, Scratch& S, double fe[8][3], double* dtElem) {
  double det, dtk;
  int status = hex8_brick_setup(in, mp, S, &det, &dtk);
  status |= hex8_brick_loop<MATSEL>(mat, mp, updHist, hist, out, S, det, dtk, fe, dtElem);
  return status;
}

// ---------------------------------------------------------------------------------------------
// C3D4: the reference's other solid element.  One Gauss point at the centroid with weight 1/6
// (GaussQuadrature3D.cpp:62-69), N = (xi, eta, zeta, 1 - xi - eta - zeta), detJ = |det| (ShapeFunction_C3D4.cpp:70).
// Same algebra as one Gauss point of the hexahedron with the edge vectors x_k - x_3 as Jacobian columns:
//   F = I + [u_k - u_3] [X_k - X_3]^-1,   f_k = (|det|/6) P grad N_k = sign(det)/6 (P cof J0)[:, k],  f_3 = -sum.
// Characteristic length = smallest altitude of the current configuration (CalculateCharacteristicLength_C3D4.cpp:5-47).
template <int MATSEL, bool WITH_DT, class Hist, class Out>
FTB_HD int tet4_element(const double X[4][3], const double U[4][3], int mat, const double* __restrict__ mp, const bool updHist,
                        const Hist& hist, const Out& out, double fe[4][3], double* dtElem) {
  if (MATSEL >= 0) mat = MATSEL;
  int status = 0;
  double J0[3][3], Uh[3][3], cJ[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      J0[i][k] = X[k][i] - X[3][i];
      Uh[i][k] = U[k][i] - U[3][i];
    }
  cofactor3(J0, cJ);
  const double det = J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2];
  if (det == 0.0 || !(det == det)) status |= 2;
  const double rdet = ftb_rcp(det);
  double F[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      F[i][j] = (Uh[i][0] * cJ[j][0] + Uh[i][1] * cJ[j][1] + Uh[i][2] * cJ[j][2]) * rdet + (i == j ? 1.0 : 0.0);
  double cF[3][3];
  cofactor3(F, cF);
  const double J = F[0][0] * cF[0][0] + F[0][1] * cF[0][1] + F[0][2] * cF[0][2];
  if (!(J > 0.0) && mat != 0) status |= 4;
  double P[3][3], Sv[6];
  GpHistory h;
  if (mat == 5) hist.load(0, h);
  status |= material_P<Out::want_S>(mat, F, cF, J, mp, &h, updHist, P, Sv);
  if (mat == 5 && updHist) hist.store(0, h);
  if (Out::enabled) out.put(0, F, J, Sv);
  const double w = (det > 0.0 ? 1.0 : -1.0) / 6.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double Q0 = (P[i][0] * cJ[0][0] + P[i][1] * cJ[1][0] + P[i][2] * cJ[2][0]) * w;
    const double Q1 = (P[i][0] * cJ[0][1] + P[i][1] * cJ[1][1] + P[i][2] * cJ[2][1]) * w;
    const double Q2 = (P[i][0] * cJ[0][2] + P[i][1] * cJ[1][2] + P[i][2] * cJ[2][2]) * w;
    fe[0][i] = Q0; fe[1][i] = Q1; fe[2][i] = Q2;
    fe[3][i] = -(Q0 + Q1 + Q2);
  }
  if (WITH_DT) {
    double x[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) x[k][c] = X[k][c] + U[k][c];
    double altMin = 1e6;  // the reference's start value (:15)
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // vertex i against the plane of the other three, index table of :6
      const double* x0 = x[i]; const double* x1 = x[(i + 1) & 3]; const double* x2 = x[(i + 2) & 3]; const double* x3 = x[(i + 3) & 3];
      const double v1[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
      const double v2[3] = {x3[0] - x1[0], x3[1] - x1[1], x3[2] - x1[2]};
      const double n[3] = {v1[1] * v2[2] - v1[2] * v2[1], -v1[0] * v2[2] + v1[2] * v2[0], v1[0] * v2[1] - v1[1] * v2[0]};
      const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      const double alt = fabs(n[0] * (x0[0] - x1[0]) + n[1] * (x0[1] - x1[1]) + n[2] * (x0[2] - x1[2])) / nn;
      if (alt < altMin) altMin = alt;
    }
    *dtElem = altMin / mp[MP_CE];
  }
  return status;
}
// lumped nodal mass of a tetrahedron: rho (1/6) |det| N_k sum_m N_m with N = 1/4 at the Gauss point (Mass3D.cpp:5-67);
// returns |det| (the reference's detJacobian)
FTB_HD double tet4_lumped_mass(const double X[4][3], const double rho, double me[4]) {
  double J0[3][3], cJ[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) J0[i][k] = X[k][i] - X[3][i];
  cofactor3(J0, cJ);
  const double det = fabs(J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2]);
  // Me[n][m] = (N_n N_m) (w detJ) rho, row sum over m in ascending order (the reference's lumping order)
  const double entry = (0.25 * 0.25) * ((1.0 / 6.0) * det) * rho;
  const double row = ((entry + entry) + entry) + entry;
#pragma unroll
  for (int k = 0; k < 4; ++k) me[k] = row;
  return det;
}
FTB_HD double tet4_volume(const double X[4][3]) {  // Geometry.cpp:66-76
  double a[3], b[3], c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { a[i] = X[0][i] - X[3][i]; b[i] = X[1][i] - X[3][i]; c[i] = X[2][i] - X[3][i]; }
  return fabs(tp3(a, b, c)) / 6.0;
}

// Lumped nodal masses of one element (src/fem/Mass/Mass3D.cpp:5-67,127-151):
// m_k = rho * sum_gp w detJ0 N_k (sum_m N_m); the partition of unity makes the
// inner sum 1 (the reference carries it numerically; difference <= 1 ulp).
// Also returns the smallest reference-configuration detJ0 (ShapeFunction_C3D8.cpp:98).
FTB_HD double hex8_lumped_mass(const double X[8][3], const double rho, double me[8]) {
  double gX[7][3];
  {
    double n[8], g[7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) n[k] = X[k][c];
      hex_modes(n, g);
#pragma unroll
      for (int m = 0; m < 7; ++m) gX[m][c] = g[m];
    }
  }
  const double a = FTB_GP_A, a2 = FTB_GP_A * FTB_GP_A;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gX[3][c] *= a; gX[4][c] *= a; gX[5][c] *= a; gX[6][c] *= a2;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) me[k] = 0.0;
  double detMin = 1e300;
  const double ks1[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
  const double ks2[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  const double ks3[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
#pragma unroll
  for (int gp = 0; gp < 8; ++gp) {
    const double s1 = FTB_GP_S1(gp), s2 = FTB_GP_S2(gp), s3 = FTB_GP_S3(gp);
    double J0[3][3], cJ[3][3];
    gp_jacobian(gX, s1, s2, s3, J0);
    cofactor3(J0, cJ);
    const double detJ = (J0[0][0] * cJ[0][0] + J0[0][1] * cJ[0][1] + J0[0][2] * cJ[0][2]) / 512.0;
    if (detJ < detMin) detMin = detJ;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double Nk = ((1.0 + ks1[k] * s1 * a) * (1.0 + ks2[k] * s2 * a) * (1.0 + ks3[k] * s3 * a)) / 8.0;
      me[k] += Nk * detJ;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) me[k] *= rho;
  return detMin;
}

}  // namespace ftb
