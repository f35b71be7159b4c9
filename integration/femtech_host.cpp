// femtech_host.cpp -- the reference-facing C++ host layer of femtech_b200.
//
// Exports the SAME C++ symbols the reference's libFemTech.a exports for the explicit-dynamics hot path
// (include/FemTech.h:22-129) and defines the global arrays those translation units define
// (src/fem/ShapeFunctions/ShapeFunctions.cpp:7-30, src/fem/AllocateArrays.cpp:5-27), so that the
// reference's shipped drivers (examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp, examples/ex9/ex9.cpp)
// compile and link UNCHANGED: put this object before libFemTech.a on the link line and add -lftb200.
// Every function is a thin marshalling shim over the C-ABI in include/ftb200.h -- no numerics here.
//
// Compiled against the reference's own headers (-I<FemTech>/include); see integration/build_dropin.sh
// and INTEGRATION.md.  Replaced translation units of the reference:
//   src/fem/ShapeFunctions/ShapeFunctions.cpp   ShapeFunctions()
//   src/fem/Mass/Mass3D.cpp                     AssembleLumpedMass() (+ aborting stubs for the implicit-only entries)
//   src/fem/AllocateArrays.cpp                  AllocateArrays()
//   src/io/output/FreeArrays.cpp                FreeArrays()
//   src/fem/SolidMechanics/GetForce.cpp, GetForce_3D.cpp, CalculateAcclerations.cpp, CheckEnergy.cpp
//   src/timestep/StableTimeStep.cpp             StableTimeStep()
//   src/fem/Solver/ExplicitDynamics.cpp         ExplicitDynamics() (a stub in the reference)
//   src/elements/ElementCalculations/CalculateStrain.cpp  CalculateStrain() (lazy F/pk2/Eavg materialisation)
// Modes: the per-call functions are the LEGACY mode -- host arrays cross PCIe on every call, the
// drivers' own host loops stay as they are; ExplicitDynamics() is the RESIDENT mode.
#include "FemTech.h"
#include "utilities.h"

#include "ftb200.h"
#include "femtech_b200_ext.h"

#include <string>

/* ---- globals defined by the replaced translation units ---------------------------------------- */
int *gptr, *dsptr, *GaussPoints, *fptr, *nShapeFunctions, *pk2ptr, *detFptr, *InternalsPtr, *gpPtr;
double *shp, *dshp, *F, *detF, *invF, *pk2, *internals, *detJacobian, *gaussWeights, *fintGQ, *B;
double *Hn_1, *Hn_2, *S0n;
double *displacements, *velocities, *velocities_half, *accelerations, *Eavg, *fe, *fe_prev, *fi, *fi_prev, *f_net;
double *fr_curr, *fr_prev, *fi_curr, *f_damp_prev, *f_damp_curr, *displacements_prev, *accelerations_prev, *stepTime;
double *mat1, *mat2, *mat3, *mat4;
int *boundary;
FILE *energyFile;

static ftb200_ctx *g_ctx = NULL;
static int *g_bc_kind = NULL;
static double g_bc_rate[4] = {0, 0, 0, 0};
static int g_energy_every = 1;

static void check(int rc) {
  if (rc) {
    FILE_LOG_SINGLE(ERROR, "femtech_b200: %s", ftb200_last_error(g_ctx));
    TerminateFemTech(rc == FTB200_ERR_CUDA ? 3 : rc);  /* same abort codes as the reference */
  }
}

static void ensure_ctx() {
  if (g_ctx) return;
  if (ndim != 3) { FILE_LOG_SINGLE(ERROR, "femtech_b200 supports ndim == 3 only"); TerminateFemTech(3); }
  for (int e = 0; e < nelements; ++e) {  /* the solid elements of the reference: C3D8 and C3D4 (ShapeFunctions.cpp:71-164) */
    const bool hex = strcmp(ElementType[e], "C3D8") == 0 && eptr[e + 1] - eptr[e] == 8;
    const bool tet = strcmp(ElementType[e], "C3D4") == 0 && eptr[e + 1] - eptr[e] == 4;
    if (!hex && !tet) {
      FILE_LOG_SINGLE(ERROR, "femtech_b200 replaces the C3D8 / C3D4 path; element %d is %s with %d nodes", e, ElementType[e],
                      eptr[e + 1] - eptr[e]);
      TerminateFemTech(3);
    }
  }
  /* one rank per GPU; with fewer GPUs than ranks (a test box) the ranks share the devices round robin */
  const int ndev = ftb200_device_count();
  const char *ev = getenv("FTB200_DEVICE");
  check(ftb200_create(world_rank, world_size, ev ? atoi(ev) : (ndev > 0 ? world_rank % ndev : 0), &g_ctx));
  check(ftb200_upload_mesh_mixed(g_ctx, coordinates, connectivity, eptr, pid, nNodes, nelements));
  check(ftb200_upload_materials(g_ctx, materialID, properties, nPIDglobal));
  check(ftb200_upload_comm(g_ctx, sendProcessCount, sendProcessID, sendNeighbourCountCum, sendNodeIndex));
}

/* ---- the shared-node exchange of the reference, kept as it is: sendNodeDisplacement -> neighbours -> recvNodeDisplacement
 * with one MPI_Isend / MPI_Irecv per neighbour (GetForce_3D.cpp:63-91, Mass3D.cpp:86-114).  The device packs and adds
 * (ftb200_*_host entry points); the transport is the host's MPI. ------------------------------------------------------ */
static void exchange_shared_nodes(int tag) {
  if (world_size == 1 || sendProcessCount == 0) return;
  MPI_Request *req = (MPI_Request *)malloc(sizeof(MPI_Request) * 2 * sendProcessCount);
  for (int i = 0; i < sendProcessCount; ++i) {
    const int location = sendNeighbourCountCum[i] * ndim, size = ndim * sendNeighbourCount[i];
    MPI_Isend(&sendNodeDisplacement[location], size, MPI_DOUBLE, sendProcessID[i], tag, MPI_COMM_WORLD, &req[i]);
  }
  for (int i = 0; i < sendProcessCount; ++i) {
    const int location = sendNeighbourCountCum[i] * ndim, size = ndim * sendNeighbourCount[i];
    MPI_Irecv(&recvNodeDisplacement[location], size, MPI_DOUBLE, sendProcessID[i], tag, MPI_COMM_WORLD, &req[sendProcessCount + i]);
  }
  MPI_Status status;
  for (int i = 0; i < 2 * sendProcessCount; ++i) MPI_Wait(&req[i], &status);
  free(req);
}

/* ---- src/fem/AllocateArrays.cpp:29-153 ----------------------------------------------------------- */
static double *zalloc(size_t n, const char *what) {
  double *p = (double *)calloc(n ? n : 1, sizeof(double));
  if (!p) { FILE_LOG_SINGLE(ERROR, "Error in allocating %s array", what); TerminateFemTech(12); }
  return p;
}
void AllocateArrays() {
  displacements = zalloc(nDOF, "displacements");
  boundary = (int *)calloc(nDOF, sizeof(int));
  if (!boundary) { FILE_LOG_SINGLE(ERROR, "Error in allocating boundary array"); TerminateFemTech(12); }
  if (ImplicitDynamic || ExplicitDynamic) {
    velocities = zalloc(nDOF, "velocities");
    accelerations = zalloc(nDOF, "accelerations");
    velocities_half = zalloc(nDOF, "velocities_half");
    displacements_prev = zalloc(nDOF, "displacements_prev");
    accelerations_prev = zalloc(nDOF, "accelerations_prev");
    Eavg = zalloc((size_t)nelements * ndim * ndim, "Eavg");
    fe = zalloc(nDOF, "fe"); fe_prev = zalloc(nDOF, "fe_prev");
    fi = zalloc(nDOF, "fi"); fi_prev = zalloc(nDOF, "fi_prev");
    f_net = zalloc(nDOF, "f_net");
    fr_prev = zalloc(nDOF, "fr_prev"); fr_curr = zalloc(nDOF, "fr_curr"); fi_curr = zalloc(nDOF, "fi_curr");
    f_damp_curr = zalloc(nDOF, "f_damp_curr"); f_damp_prev = zalloc(nDOF, "f_damp_prev");
    std::string energyFileName = "energy_" + uid + ".dat";
    energyFile = fopen(energyFileName.c_str(), "w");
    fprintf(energyFile, "# Energy for FEM\n");
    fprintf(energyFile, "# Time  Winternal   Wexternal   WKE   total\n");
    stepTime = (double *)malloc((MAXPLOTSTEPS) * sizeof(double));
    if (!stepTime) { FILE_LOG_SINGLE(ERROR, "Error in allocating stepTime array"); TerminateFemTech(12); }
  }
}

/* ---- src/fem/ShapeFunctions/ShapeFunctions.cpp:32-255: only the offset tables are kept on the host; the
 * per-Gauss-point tables (shp, dshp, detJacobian: 3.8 kB per element) are never built -- the device
 * recomputes them.  F, detF and pk2 are materialised lazily by CalculateStrain(). ------------------------- */
void ShapeFunctions() {
  ensure_ctx();
  GaussPoints = (int *)malloc(nelements * sizeof(int));
  nShapeFunctions = (int *)malloc(nelements * sizeof(int));
  gptr = (int *)malloc((nelements + 1) * sizeof(int)); dsptr = (int *)malloc((nelements + 1) * sizeof(int));
  gpPtr = (int *)malloc((nelements + 1) * sizeof(int)); fptr = (int *)malloc((nelements + 1) * sizeof(int));
  pk2ptr = (int *)malloc((nelements + 1) * sizeof(int)); detFptr = (int *)malloc((nelements + 1) * sizeof(int));
  InternalsPtr = (int *)malloc((nelements + 1) * sizeof(int));
  int ngp = 0, nshp = 0;  /* offset tables of ShapeFunctions.cpp:71-164: 8 points x 8 functions (C3D8), 1 x 4 (C3D4) */
  for (int i = 0; i <= nelements; ++i) {
    gptr[i] = nshp; dsptr[i] = 3 * nshp; gpPtr[i] = ngp; fptr[i] = 9 * ngp; pk2ptr[i] = 6 * ngp; detFptr[i] = ngp;
    InternalsPtr[i] = MAXINTERNALVARS * ngp;
    if (i < nelements) {
      const bool tet = eptr[i + 1] - eptr[i] == 4;
      GaussPoints[i] = tet ? 1 : 8; nShapeFunctions[i] = tet ? 4 : 8;
      ngp += GaussPoints[i]; nshp += GaussPoints[i] * nShapeFunctions[i];
    }
  }
  double minDetJ = 0;
  check(ftb200_shape_functions(g_ctx, &minDetJ));
}

/* ---- src/fem/Mass/Mass3D.cpp:127-157 ---------------------------------------------------------------- */
void AssembleLumpedMass(void) {
  ensure_ctx();
  mass = (double *)calloc(nDOF, sizeof(double));
  if (!mass) { FILE_LOG_SINGLE(ERROR, "Allocation of mass matrix failed"); TerminateFemTech(12); }
  if (world_size == 1) {
    check(ftb200_lumped_mass(g_ctx, mass));
    return;
  }
  /* Mass3D.cpp:77-125 updateMassMatrixNeighbour: the shared nodes receive their neighbours' share */
  check(ftb200_lumped_mass(g_ctx, NULL));
  check(ftb200_halo_pack_host(g_ctx, 1, sendNodeDisplacement));
  exchange_shared_nodes(7132);
  check(ftb200_halo_add_host(g_ctx, 1, recvNodeDisplacement));
  check(ftb200_get_mass(g_ctx, mass));
}
static void implicit_only(const char *what) {
  FILE_LOG_SINGLE(ERROR, "%s belongs to the dense implicit solvers, which femtech_b200 does not replace", what);
  TerminateFemTech(3);
}
void MassElementMatrix(double *, int) { implicit_only("MassElementMatrix"); }
void LumpMassMatrix(void) { implicit_only("LumpMassMatrix"); }
void updateMassMatrixNeighbour(void) { implicit_only("updateMassMatrixNeighbour"); }

/* ---- src/fem/SolidMechanics/GetForce.cpp:12-25, GetForce_3D.cpp:5-53 ----------------------------------- */
static unsigned long long g_state_gen = 1;   /* bumped whenever the device displacements change */
void GetForce_3D() {
  g_state_gen++;
  bool anyFe = false;
  for (int i = 0; i < nDOF && !anyFe; ++i) anyFe = fe[i] != 0.0;
  if (world_size == 1) {
    check(ftb200_get_force(g_ctx, displacements, anyFe ? fe : NULL, dt, fi, f_net));
    return;
  }
  /* GetForce_3D.cpp:54-102 updateInternalForceNeighbour around the reference's own MPI exchange.  fe is uploaded on every
   * rank as soon as one rank has any (the captured kernels' arguments must agree across the step). */
  int anyGlobal = anyFe ? 1 : 0;
  MPI_Allreduce(MPI_IN_PLACE, &anyGlobal, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  check(ftb200_get_force_begin(g_ctx, displacements, anyGlobal ? fe : NULL, dt, sendNodeDisplacement));
  exchange_shared_nodes(2169);
  check(ftb200_get_force_end(g_ctx, recvNodeDisplacement, fi, f_net));
}
void GetForce() {
  if (ndim != 3) { FILE_LOG_SINGLE(ERROR, "GetForce function not yet implemented for %dD", ndim); TerminateFemTech(3); }
  GetForce_3D();
}
/* ---- src/fem/SolidMechanics/CalculateAcclerations.cpp:4-13 ------------------------------------------- */
void CalculateAccelerations() { check(ftb200_calculate_accelerations(g_ctx, boundary, accelerations)); }

/* ---- src/timestep/StableTimeStep.cpp:4-40 ------------------------------------------------------------- */
double StableTimeStep() {
  double dtMin = huge;
  check(ftb200_stable_time_step(g_ctx, displacements, boundary, &dtMin));
  MPI_Allreduce(MPI_IN_PLACE, &dtMin, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);
  if (dtMin < FailureTimeStep) {
    FILE_LOG_MASTER(ERROR, "Timestep too small, dt = %15.9e", dtMin);
    TerminateFemTech(19);
  }
  return dtMin;
}

/* ---- src/fem/SolidMechanics/CheckEnergy.cpp:3-85 ------------------------------------------------------- */
void CheckEnergy(double time, int writeFlag) {
  static double Wint_n = 0.0, Wext_n = 0.0;
  double part[3];
  check(ftb200_check_energy(g_ctx, displacements, displacements_prev, velocities, accelerations, accelerations_prev, fi,
                            fi_prev, fe, fe_prev, boundary, part));
  double WKE_Total = 0.0, Wint_n_total = 0.0, Wext_n_total = 0.0;
  MPI_Reduce(&part[0], &WKE_Total, 1, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
  MPI_Reduce(&part[1], &Wint_n_total, 1, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
  MPI_Reduce(&part[2], &Wext_n_total, 1, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
  if (world_rank == 0) {
    Wint_n += Wint_n_total;
    Wext_n += Wext_n_total;
    double total = fabs(WKE_Total + Wint_n - Wext_n);
    double max = fabs(Wint_n);
    if (max < fabs(Wext_n)) max = fabs(Wext_n);
    if (max < fabs(WKE_Total)) max = fabs(WKE_Total);
    if (total > 0.01 * max)
      FILE_LOG_MASTER(WARNING, "Energy Violation. Total = %15.9e, Max = %15.9e, Error%% : %10.2f", total, max, total * 100.0 / max);
    if (writeFlag == 0) fprintf(energyFile, "%12.6e %12.6e  %12.6e  %12.6e %12.6e\n", time, Wint_n, Wext_n, WKE_Total, total);
  }
}

/* ---- src/elements/ElementCalculations/CalculateStrain.cpp:77-97 + lazy Gauss-point outputs --------------- */
void CalculateStrain() {
  if (!F) F = (double *)calloc((size_t)fptr[nelements] + 1, sizeof(double));
  if (!detF) detF = (double *)calloc((size_t)detFptr[nelements] + 1, sizeof(double));
  if (!pk2) pk2 = (double *)calloc((size_t)pk2ptr[nelements] + 1, sizeof(double));
  check(ftb200_get_gp_outputs(g_ctx, F, detF, pk2, Eavg));
}

/* ---- CalculateStrain.cpp:8-75.  The brain drivers call this per element right after GetForce (ex5.cpp:1313-1317,
 * :556); the device evaluates all elements in one launch and the per-element calls read the cached arrays until
 * the next force evaluation.  Like the reference it leaves E of the element in Eavg. -------------------------- */
static double *g_ps_max = NULL, *g_ps_min = NULL, *g_ps_shear = NULL;
static unsigned long long g_ps_gen = 0;
void CalculateMaximumPrincipalStrain(int elm, double *currentStrainMax, double *currentStrainMin, double *currentShearMax) {
  if (g_ps_gen != g_state_gen) {
    if (!g_ps_max) {
      g_ps_max = (double *)malloc(sizeof(double) * nelements);
      g_ps_min = (double *)malloc(sizeof(double) * nelements);
      g_ps_shear = (double *)malloc(sizeof(double) * nelements);
    }
    check(ftb200_principal_strains(g_ctx, g_ps_max, g_ps_min, g_ps_shear, NULL));
    check(ftb200_get_gp_outputs(g_ctx, NULL, NULL, NULL, Eavg));
    g_ps_gen = g_state_gen;
  }
  *currentStrainMax = g_ps_max[elm];
  *currentStrainMin = g_ps_min[elm];
  *currentShearMax = g_ps_shear[elm];
}

/* Resident counterpart of InitInjuryCriterion / CalculateInjuryCriterions (ex5.cpp:1251-1430), femtech_b200_ext.h */
void femtech_b200_injury_begin(const int *injuryExcludePID, int injuryExcludePIDCount) {
  ensure_ctx();
  check(ftb200_injury_begin(g_ctx, injuryExcludePID, injuryExcludePIDCount, NULL));
}
void femtech_b200_injury_results(double scalars12[12], int extreme_elems4[4], unsigned char *flags, double *PS_Old,
                                 double *PSxSRArray, double volumes5[5]) {
  check(ftb200_injury_get(g_ctx, scalars12, extreme_elems4, flags, PS_Old, PSxSRArray, volumes5));
}

/* ---- include/FemTech.h:50 -- a stub in the reference (src/fem/Solver/ExplicitDynamics.cpp:7-9); here the whole
 * loop of the drivers (Benchmarking-Parallel.cpp:83-171) resident on the GPU.  The driver's boundary-condition
 * callback becomes the descriptor set with femtech_b200_set_bc() (femtech_b200_ext.h). ---------------------- */
void femtech_b200_set_bc(const int *bc_kind, const double bc_rate[4], int energy_every) {
  if (!g_bc_kind) g_bc_kind = (int *)malloc(sizeof(int) * nDOF);
  memcpy(g_bc_kind, bc_kind, sizeof(int) * nDOF);
  memcpy(g_bc_rate, bc_rate, sizeof(g_bc_rate));
  g_energy_every = energy_every;
}
/* Resident counterpart of InitBoundaryCondition / ApplyAccBoundaryConditions (ex5.cpp:574-912, :339-371) */
static bool g_rigid = false;
void femtech_b200_set_rigid_bc(const int sizes[6], const double *const t[6], const double *const v[6], const int *boundaryID,
                               int boundarySize, int energy_every) {
  ensure_ctx();
  check(ftb200_set_rigid_bc(g_ctx, sizes, t, v, boundaryID, boundarySize));
  g_rigid = true;
  g_energy_every = energy_every;
}
/* Several ranks: one partition per GPU, the reference's ParMETIS split (PartitionMesh.cpp:23-83) as it stands in the
 * globals.  Two transports for the shared-node sum and the dt MIN of every step:
 *   p2p   peer-memory windows over NVLink (CUDA IPC handles exchanged once with MPI_Allgather, then no MPI and no host in
 *         the loop: the whole step is a CUDA graph) -- the default when every rank has its own GPU;
 *   host  the reference's own MPI_Isend / MPI_Irecv exchange and MPI_Allreduce(MIN), one round per step, the device packing
 *         and adding -- the default when ranks share a GPU (processes on one device time-slice: a kernel that waits for a
 *         peer's flag would wait for the peer's time slice), FTB200_MPI_TRANSPORT=host|p2p overrides. */
static bool g_p2p = false, g_p2p_ready = false;
static long long g_last_steps = 0;
static double g_last_seconds = 0.0;
long long femtech_b200_last_steps(void) { return g_last_steps; }
double femtech_b200_last_seconds(void) { return g_last_seconds; }
const char *femtech_b200_last_transport(void) { return world_size == 1 ? "single" : (g_p2p ? "p2p" : "host"); }
static void setup_p2p() {
  if (g_p2p_ready) return;
  unsigned char handle[64];
  check(ftb200_p2p_export(g_ctx, handle, NULL));
  unsigned char *all = (unsigned char *)malloc((size_t)64 * world_size);
  MPI_Allgather(handle, 64, MPI_BYTE, all, 64, MPI_BYTE, MPI_COMM_WORLD);
  /* where this rank's slice sits in each neighbour's window: the neighbour tells (its cumulative count for us, our index
   * in its list, its total) -- the send lists are symmetric (PartitionMesh.cpp:566-1128) */
  const int nb = sendProcessCount;
  int *mine = (int *)malloc(sizeof(int) * 3 * (nb + 1)), *theirs = (int *)malloc(sizeof(int) * 3 * (nb + 1));
  MPI_Request *req = (MPI_Request *)malloc(sizeof(MPI_Request) * 2 * (nb + 1));
  for (int i = 0; i < nb; ++i) {
    mine[3 * i] = sendNeighbourCountCum[i]; mine[3 * i + 1] = i; mine[3 * i + 2] = sendNeighbourCountCum[nb];
    MPI_Isend(&mine[3 * i], 3, MPI_INT, sendProcessID[i], 2170, MPI_COMM_WORLD, &req[i]);
    MPI_Irecv(&theirs[3 * i], 3, MPI_INT, sendProcessID[i], 2170, MPI_COMM_WORLD, &req[nb + i]);
  }
  MPI_Status status;
  for (int i = 0; i < 2 * nb; ++i) MPI_Wait(&req[i], &status);
  int *off = (int *)malloc(sizeof(int) * (nb + 1)), *idx = (int *)malloc(sizeof(int) * (nb + 1)), *tot = (int *)malloc(sizeof(int) * (nb + 1));
  for (int i = 0; i < nb; ++i) { off[i] = theirs[3 * i]; idx[i] = theirs[3 * i + 1]; tot[i] = theirs[3 * i + 2]; }
  check(ftb200_p2p_import(g_ctx, all, 0, off, idx, tot));
  MPI_Barrier(MPI_COMM_WORLD);
  free(all); free(mine); free(theirs); free(req); free(off); free(idx); free(tot);
  g_p2p_ready = true;
}
/* `n` steps of the loop (fewer if timeFinal is reached: the remaining iterations are no-ops on the device) */
static void run_steps(double timeFinal, long long n) {
  if (world_size == 1 || g_p2p) {
    check(ftb200_explicit_run_async(g_ctx, timeFinal, n));
    return;
  }
  check(ftb200_run_begin(g_ctx, timeFinal, n));
  for (long long i = 0; i < n; ++i) {
    double dtl = 0.0;
    check(ftb200_step_begin_host(g_ctx, sendNodeDisplacement, &dtl));
    exchange_shared_nodes(2169);
    MPI_Allreduce(MPI_IN_PLACE, &dtl, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);  /* StableTimeStep.cpp:33 */
    check(ftb200_step_end_host(g_ctx, recvNodeDisplacement, dtl));
  }
}
void ExplicitDynamics(double timeFinal, char *name) {
  (void)name;
  ensure_ctx();
  g_state_gen++;
  const double wall0 = MPI_Wtime();
  if (!g_bc_kind && g_rigid) {  /* only the rigid-body condition: no other dof is prescribed */
    g_bc_kind = (int *)calloc(nDOF, sizeof(int));
  }
  if (!g_bc_kind) { FILE_LOG_SINGLE(ERROR, "ExplicitDynamics: call femtech_b200_set_bc() or femtech_b200_set_rigid_bc() first"); TerminateFemTech(3); }
  check(ftb200_set_state(g_ctx, displacements, velocities, accelerations, boundary));
  check(ftb200_set_bc(g_ctx, g_bc_kind, g_bc_rate));
  if (world_size == 1) {
    check(ftb200_explicit_begin(g_ctx, Time, ExplicitTimeStepReduction, FailureTimeStep, g_energy_every));
  } else {
    const char *tr = getenv("FTB200_MPI_TRANSPORT");
    g_p2p = tr ? strcmp(tr, "p2p") == 0 : ftb200_device_count() >= world_size;
    /* step 0 of the drivers (Benchmarking-Parallel.cpp:83-91) with the two cross-rank operations on the host's MPI */
    check(ftb200_explicit_begin_dt(g_ctx, Time, ExplicitTimeStepReduction, FailureTimeStep, g_energy_every, NULL));
    double dtl = 0.0;
    check(ftb200_get_dtmin(g_ctx, &dtl));
    MPI_Allreduce(MPI_IN_PLACE, &dtl, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);
    check(ftb200_set_dtmin(g_ctx, dtl));
    check(ftb200_explicit_begin_force_host(g_ctx, sendNodeDisplacement));
    exchange_shared_nodes(2169);
    check(ftb200_explicit_begin_finish_host(g_ctx, recvNodeDisplacement));
    if (g_p2p) setup_p2p();
  }
  /* The loop runs in slices of RING/2 steps; every finished step's record (Time, dt, step, status, energies) is written by
   * the device into the pinned host ring of ftb200_step_ring, and the energy file gets the line CheckEnergy would have
   * written for that step (CheckEnergy.cpp:66-83: "%12.6e %12.6e  %12.6e  %12.6e %12.6e") -- one line per step, as in the
   * reference's drivers, without a host round trip per step.  Several ranks: the records carry this rank's share of the
   * sums (owner rule of CheckEnergy.cpp:21-33 on the device); rank 0 adds the shares of a slice with one MPI_Reduce. */
  const long long RING = 512;
  double *ring = NULL;
  check(ftb200_step_ring(g_ctx, RING, &ring));
  long long steps = 0, done = 0;
  int st = 0;
  double *part = (double *)malloc(sizeof(double) * 3 * RING), *sum = (double *)malloc(sizeof(double) * 3 * RING);
  for (;;) {
    run_steps(timeFinal, RING / 2);
    long long now = 0;
    check(ftb200_explicit_poll(g_ctx, &now, &Time, &dt, &st));
    const long long cnt = now - done;
    for (long long k = done + 1; k <= now; ++k) {
      const volatile double *r = ring + 8 * ((k - 1) % RING);
      if (r[2] != (double)k) { FILE_LOG_SINGLE(ERROR, "ExplicitDynamics: step ring out of sequence at step %lld", k); TerminateFemTech(3); }
      part[3 * (k - done - 1)] = r[4]; part[3 * (k - done - 1) + 1] = r[5]; part[3 * (k - done - 1) + 2] = r[6];
    }
    if (world_size > 1) {
      if (cnt > 0) MPI_Reduce(part, sum, (int)(3 * cnt), MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
    } else {
      memcpy(sum, part, sizeof(double) * 3 * cnt);
    }
    if (g_energy_every && world_rank == 0)
      for (long long k = done + 1; k <= now; ++k) {
        const volatile double *r = ring + 8 * ((k - 1) % RING);
        const double *e = sum + 3 * (k - done - 1);
        fprintf(energyFile, "%12.6e %12.6e  %12.6e  %12.6e %12.6e\n", r[0], e[0], e[1], e[2], fabs(e[2] + e[0] - e[1]));
      }
    steps += cnt;
    if (now == done || !(Time < timeFinal) || (st & 16)) break;
    done = now;
  }
  free(part); free(sum);
  check(ftb200_step_ring(g_ctx, 0, NULL));
  if (st & 1) { FILE_LOG_SINGLE(ERROR, "Unknown material type"); TerminateFemTech(1); }
  if (st & 16) { FILE_LOG_SINGLE(ERROR, "Timestep too small, dt below FailureTimeStep"); TerminateFemTech(19); }
  check(ftb200_get_state(g_ctx, displacements, velocities, accelerations, boundary, fi, f_net));
  g_last_steps = steps;
  g_last_seconds = MPI_Wtime() - wall0;
  FILE_LOG_MASTER(INFO, "ExplicitDynamics: %lld steps on the GPU, Time = %15.6e", steps, Time);
}

/* ---- src/io/output/FreeArrays.cpp:4-83 --------------------------------------------------------------------- */
void FreeArrays() {
  if (g_ctx) { ftb200_destroy(g_ctx); g_ctx = NULL; }
  void *ptrs[] = {coordinates, connectivity, globalNodeID, pid, global_eid, eptr, shp, dshp, dsptr, gptr, nShapeFunctions, C,
                  gaussWeights, gpPtr, detJacobian, mass, stiffness, rhs, displacements, velocities, accelerations,
                  accelerations_prev, boundary, velocities_half, fe, fe_prev, fi, f_net, fr_prev, fr_curr, fi_prev, fi_curr,
                  f_damp_prev, f_damp_curr, displacements_prev, F, pk2, pk2ptr, fptr, materialID, properties, detF, invF,
                  Eavg, detFptr, InternalsPtr, internals, GaussPoints, recvNodeDisplacement, sendProcessID,
                  sendNeighbourCount, sendNeighbourCountCum, sendNodeIndex, sendNodeDisplacement, stepTime, mat1, mat2,
                  mat3, mat4, fintGQ, B, Hn_1, Hn_2, S0n, g_bc_kind, g_ps_max, g_ps_min, g_ps_shear};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free1DArray(ptrs[i]);
  g_bc_kind = NULL;
  g_ps_max = g_ps_min = g_ps_shear = NULL;
  g_ps_gen = 0;
  if (ElementType != NULL) {
    for (int i = 0; i < nelements; i++) free(ElementType[i]);
    free(ElementType);
    ElementType = NULL;
  }
  if (world_rank == 0 && energyFile) fclose(energyFile);
}
