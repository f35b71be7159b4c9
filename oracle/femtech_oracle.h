/*
 * femtech_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the FemTech explicit-dynamics hot path, used
 * as the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (femtech_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * bit-for-bit against vectors dumped from the reference's own sources compiled
 * in this container (oracle/ref/build_ref.sh -> oracle/_ref/ref_dump, recipe
 * and generating script committed; vectors under tests/golden/), and against
 * the reference's only known-answer test (examples/ex9 vs abaqus.rpt, 5 %).
 * The BLAS under the reference is external and unversioned (CMakeLists.txt:140);
 * "reference result" is defined with the naive left-to-right shim in
 * oracle/ref/blas_shim.c, and this file uses the same summation order.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef FEMTECH_ORACLE_H
#define FEMTECH_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAXMATPARAMS 9 /* include/GlobalVariables.h:6 */

/* One rank's view of the model: exactly the global arrays of
 * include/GlobalVariables.h:18-127 that the hot path touches.  All arrays are
 * caller-owned (numpy in the tests). */
typedef struct oracle_state {
  int nNodes, nElements, nPID;
  /* mesh (GlobalVariables.h:32-33,47-49) */
  const double *coordinates; /* [3*nNodes] AoS xyz */
  const int *connectivity;   /* [8*nElements] local node ids, C3D8 order */
  const int *pid;            /* [nElements] */
  const int *materialID;     /* [nPID] */
  const double *properties;  /* [9*nPID] rho mu lambda k1 k2 g1 t1 g2 t2 */
  /* per-Gauss-point tables (ShapeFunctions.cpp:181-204) */
  double *shp;          /* [64*nElements] */
  double *dshp;         /* [192*nElements] dN/dX, reference config */
  double *detJacobian;  /* [8*nElements] */
  double *gaussWeights; /* [8*nElements] */
  double *F;            /* [72*nElements] col-major 3x3 per GP */
  double *detF;         /* [8*nElements] */
  double *pk2;          /* [48*nElements] Voigt 11,22,33,23,13,12 */
  double *Hn_1, *Hn_2, *S0n; /* [72*nElements] or NULL (material 5 only) */
  /* nodal state (AllocateArrays.cpp:29-153) */
  double *displacements, *velocities, *velocities_half, *accelerations;
  double *mass, *fe, *fi, *f_net;
  double *displacements_prev, *accelerations_prev, *fi_prev, *fe_prev;
  int *boundary;
  /* communication pattern (PartitionMesh.cpp:566-1128) */
  int world_rank;
  int sendProcessCount;
  const int *sendProcessID;
  const int *sendNeighbourCountCum;
  const int *sendNodeIndex;
  /* driver globals (Benchmarking-Parallel.cpp:9-17) */
  double Time, dt;
  /* CheckEnergy function-statics (CheckEnergy.cpp:4-5) */
  double Wint_n, Wext_n;
  /* mixed C3D8/C3D4 meshes: gpoff[nElements+1] = Gauss points before each element (8 per hexahedron, 1 per
   * tetrahedron, ShapeFunctions.cpp:71-164); connectivity is then packed 8 or 4 per element (eptr) and the per-GP
   * arrays hold gpoff[nElements] points.  NULL = all hexahedra (the layouts documented above). */
  const int *gpoff;
} oracle_state;

/* ShapeFunctions.cpp:32-255 + ShapeFunction_C3D8.cpp:4-128 (hex8 only). Also
 * resets F<-I, detF<-1, pk2<-0 and zeroes the history arrays. */
void oracle_ShapeFunctions(oracle_state *s);
/* Mass3D.cpp:127-157 without the halo (the caller sums shared nodes with
 * oracle_halo_sum). mass must be zeroed by the caller. */
void oracle_AssembleLumpedMass_local(oracle_state *s);
/* GetForce_3D.cpp:11-46: f_net<-fe, fi<-0, element/GP loop, scatter. Returns
 * 0, or 1 when an element carries an unknown material (StressUpdate.cpp:24). */
int oracle_GetForce_local(oracle_state *s);
/* GetForce_3D.cpp:49-51 */
void oracle_GetForce_finish(oracle_state *s);
/* CalculateAcclerations.cpp:4-13 */
void oracle_CalculateAccelerations(oracle_state *s);
/* StableTimeStep.cpp:11-30 (no Allreduce, no abort) */
double oracle_StableTimeStep_local(const oracle_state *s);
/* CheckEnergy.cpp:19-52: partial sums out[0..2] = WKE, Wint, Wext (already
 * multiplied by 0.5) of the nodes this rank owns. */
void oracle_CheckEnergy_local(const oracle_state *s, double out[3]);
/* GetForce_3D.cpp:54-102 / Mass3D.cpp:77-125 for P ranks emulated in one
 * process: field 0 = fi, 1 = mass. */
void oracle_halo_sum(oracle_state **ranks, int nranks, int field);

/* The explicit loop of Benchmarking-Parallel.cpp:83-171 (== ex9.cpp) for P
 * emulated ranks.  Boundary conditions are a descriptor instead of the
 * driver's coordinate scan (:184-244): bc_kind[r][dof] = 0 free, k>0 -> dof is
 * prescribed with u = Time*bc_rate[k], v = bc_rate[k], a = 0, boundary = 1.
 * Runs while Time < tMax and at most maxSteps iterations.  Step 0 (BC, dt,
 * GetForce, accelerations) is done when s->Time == 0 and first_call != 0.
 * Per-step records (optional, may be NULL): dt_hist[k] = dt used by step k,
 * energy_hist[4k..] = Wint, Wext, WKE, total (rank-0 values).
 * Returns the number of loop iterations executed, or -19 when dt drops below
 * FailureTimeStep (StableTimeStep.cpp:35-38), -1 on an unknown material. */
int oracle_run_explicit(oracle_state **ranks, int nranks, int *const *bc_kind,
                        const double *bc_rate, double tMax, int maxSteps,
                        double ExplicitTimeStepReduction,
                        double FailureTimeStep, int first_call,
                        double *dt_hist, double *energy_hist);

/* Element-level helpers exported for unit tests */
double oracle_volumeHexahedron(const double *c24);
double oracle_areaHexahedronFace(const double *c24, const int *index4);
double oracle_CalculateTimeStep(const oracle_state *s, int e);
void oracle_CalculateStrain(const oracle_state *s, double *Eavg /*[9*nE]*/);

/* ---- injury criteria of the brain drivers ---------------------------------
 * Per-rank state of examples/ex5/ex5.cpp:62-83,1251-1306.  All arrays are caller-owned with capacity
 * nElements.  PS_Old is zero-initialised here (the reference mallocs it without initialising, ex5.cpp:1285; a
 * fresh glibc mmap gives zeros). */
typedef struct oracle_injury {
  int nElementsInjury;
  int *elementIDInjury;
  int *MPSgt15, *MPSgt30, *MPSRgt120, *MPSxSRgt28;
  double *PS_Old, *PSxSRArray;
  int *maxElemListMPS95, *maxElemListMPSxSR95;
  int maxElemCountMPS95, maxElemCountMPSxSR95;
  double maxStrain, minStrain, maxShear, maxPSxSR;
  int maxElem, minElem, shearElem, maxElemPSxSR;
  double maxT, minT, maxShearT, maxTimePSxSR;
  double maxMPS95, maxTimeMPS95, maxMPSxSR95, maxTimeMPSxSR95;
} oracle_injury;

/* CalculateStrain.cpp:8-75: E = mean over GP of 0.5 (F^T F - I) (written to Eavg_e[9] when not NULL), closed-form
 * eigenvalues, max clipped at >= 0, min at <= 0, shear = (max - min)/2 of the unclipped values. */
void oracle_CalculateMaximumPrincipalStrain(const oracle_state *s, int elm, double *Eavg_e, double *currentStrainMax,
                                            double *currentStrainMin, double *currentShearMax);
/* ex5.cpp:1251-1306 */
void oracle_InitInjuryCriterion(const oracle_state *s, oracle_injury *inj, const int *injuryExcludePID,
                                int injuryExcludePIDCount);
/* math.cpp:160-199 (what math.cpp:235-332 always reduces to: its refinement loop never runs since N == totalSize):
 * element (int)(0.95*total) - 1 of the ascending union of all ranks' arrays. */
double oracle_compute95thPercentileValue(double *const *data, const int *sizes, int nranks);
/* ex5.cpp:1311-1430 for P emulated ranks (Time, dt = the driver globals at the call, ex5.cpp:240) */
void oracle_CalculateInjuryCriterions(oracle_state **ranks, oracle_injury **inj, int nranks, double Time, double dt);
/* ex5.cpp:1049-1066 + Elements.cpp:30-38: reference-configuration volumes out[0..3] = MPS>15 %, MPS>30 % (within
 * the >15 % set), MPSR>120, MPSxSR>28, out[4] = volume of all included elements (this rank's share). */
void oracle_injury_volumes(const oracle_state *s, const oracle_injury *inj, double out[5]);
/* oracle_run_explicit with CalculateInjuryCriterions after CheckEnergy of every step (ex5.cpp:237-240); inj may
 * be NULL. */
int oracle_run_explicit_injury(oracle_state **ranks, int nranks, int *const *bc_kind, const double *bc_rate,
                               double tMax, int maxSteps, double ExplicitTimeStepReduction, double FailureTimeStep,
                               int first_call, double *dt_hist, double *energy_hist, oracle_injury **inj);

/* ---- rigid-body prescribed-motion boundary condition of the brain drivers ----
 * examples/ex5/ex5.cpp:339-371 (ApplyAccBoundaryConditions), :574-912 (InitBoundaryCondition), :976-1020
 * (computeDerivatives), src/math/math.cpp:99-158 (interpolateLinear, quaternions).  The driver integrates 12 states
 * y = [omega(3), r(3) = generator of the rotation quaternion, v(3), d(3)] with boost::numeric::odeint
 * runge_kutta_dopri5::do_step(sys, y, ydot, Time - dt, dt) (FSAL form: ydot in = derivative at t, out = at t + dt).
 * Boost is NOT in this image and not vendored by the reference (third-party/boost_1_71_0.zip is a missing blob):
 * oracle_dopri5_step restates the published Dormand-Prince 5(4) tableau in Boost's evaluation order; parity of that
 * one function is UNPINNED, everything around it is pinned against the reference library (quaternionExp,
 * quaternionInverse, quaternionRotate, crossProduct, interpolateLinear via oracle/ref/ref_dump.cpp). */
typedef struct oracle_rigid {
  /* acceleration time traces, seconds / (rad/s^2 | m/s^2): index 0..2 angular x,y,z, 3..5 linear x,y,z */
  int size[6];
  const double *t[6];
  const double *v[6];
  double y[12], ydot[12];   /* yInt, ydotInt (ex5.cpp:105) */
  int boundarySize;
  int *boundaryID;          /* caller-owned, capacity nNodes */
} oracle_rigid;
double oracle_interpolateLinear(int n, const double *x, const double *y, double value);
void oracle_quaternionExp(const double *q1, double *q2);
void oracle_quaternionRotate(const double *v, const double *R, const double *Rinv, double *vp);
void oracle_computeDerivatives(const oracle_rigid *rb, const double *y, double *ydot, double t);
void oracle_dopri5_step(const oracle_rigid *rb, double *y, double *ydot, double t, double dt);
/* ex5.cpp:819-911: boundary nodes = nodes of the elements whose part has material 0 (single rank: no neighbour
 * exchange), all three dofs constrained, u = v = a = 0, y = ydot = 0 */
void oracle_InitRigidBoundary(oracle_state *s, oracle_rigid *rb);
/* ex5.cpp:339-371 with Time, dt of the driver */
void oracle_ApplyAccBoundaryConditions(oracle_state *s, oracle_rigid *rb, double Time, double dt);
/* The ex5 time loop (ex5.cpp:186-295): as oracle_run_explicit_injury with ApplyAccBoundaryConditions in place of
 * the benchmark BC; single rank. */
int oracle_run_explicit_rigid(oracle_state *s, oracle_rigid *rb, double tMax, int maxSteps,
                              double ExplicitTimeStepReduction, double FailureTimeStep, int first_call, double *dt_hist,
                              double *energy_hist, oracle_injury *inj);

#ifdef __cplusplus
}
#endif
#endif
