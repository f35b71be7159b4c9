/*
 * ref_dump.cpp -- oracle harness driver.  TEST INFRASTRUCTURE ONLY.
 *
 * Links against the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by build_ref.sh) and drives them through the reference's own
 * public API (include/FemTech.h) in the call order of the shipped explicit
 * drivers (examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:19-182 and
 * examples/ex9/ex9.cpp, whose time loop lives in main()).  It adds what those
 * drivers lack: a step limit, full-precision binary dumps of every array on
 * the hot path, and a per-step dt record.  The time loop and the boundary
 * condition below restate the driver (:83-171, :184-244); all numerics are the
 * reference library's.
 *
 * With REF_INJURY_EXCLUDE set in the environment ("" or a comma-separated pid list) the loop also evaluates the
 * brain drivers' injury criteria after CheckEnergy of every step, in the call order of examples/ex5/ex5.cpp:237-240:
 * the driver-side bookkeeping (ex5.cpp:1251-1430, not part of the library) is restated below around the
 * reference library's own CalculateMaximumPrincipalStrain, compute95thPercentileValue and computePartVolume.
 *
 * usage: ref_dump <mesh.inp|.k> <out-prefix> <maxSteps> <tMax> <dMax> [cubeL] [warmupSteps] [nodump]
 *        (materials.dat is read from the current directory, ReadMaterials.cpp:11)
 *        warmupSteps: loop iterations excluded from the reported loop time (bench.py);
 *        nodump: write only scalars (timing runs on larger meshes)
 * output: <out-prefix>.rank<r>.bin  -- records: name[32] dtype[8] count(int64) payload
 */
#include "FemTech.h"

#include <stdint.h>
#include <sys/time.h>
#include <vector>

double Time, dt;
int nSteps;
double ExplicitTimeStepReduction = 0.8;
double FailureTimeStep = 1e-11;
int nPlotSteps = 50;
bool ImplicitStatic = false;
bool ImplicitDynamic = false;
bool ExplicitDynamic = true;

static double g_cubeL = 0.005;
static FILE *g_out = NULL;

static bool g_nodump = false;
static void put(const char *name, const char *dtype, const void *p, int64_t count, size_t esz) {
  if (g_nodump && count > 4) return;
  char nb[32] = {0}, tb[8] = {0};
  strncpy(nb, name, 31);
  strncpy(tb, dtype, 7);
  fwrite(nb, 1, 32, g_out);
  fwrite(tb, 1, 8, g_out);
  fwrite(&count, 8, 1, g_out);
  if (count > 0) fwrite(p, esz, (size_t)count, g_out);
}
static void putd(const char *name, const double *p, int64_t n) { put(name, "f8", p, p ? n : 0, 8); }
static void puti(const char *name, const int *p, int64_t n) { put(name, "i4", p, p ? n : 0, 4); }
static void puts1(const char *name, double v) { putd(name, &v, 1); }
static void puti1(const char *name, int v) { puti(name, &v, 1); }

/* ---- injury criteria bookkeeping of ex5.cpp:62-83,1251-1430 (driver code in the reference, restated) ---- */
struct Injury {
  bool on = false;
  std::vector<int> excl, elems, gt15, gt30, r120, xsr28, list95, listx95;
  std::vector<double> psOld, psxsr, hist95, histx95;
  double maxStrain = 0, minStrain = 0, maxShear = 0, maxPSxSR = 0, maxT = 0, minT = 0, maxShearT = 0, maxTimePSxSR = 0;
  int maxElem = 0, minElem = 0, shearElem = 0, maxElemPSxSR = 0;
  double max95 = 0, t95 = 0, maxx95 = 0, tx95 = 0;
} g_inj;

static void InjuryInit(const char *spec) {
  g_inj.on = true;
  for (const char *p = spec; *p;) {
    char *end;
    long v = strtol(p, &end, 10);
    if (end == p) break;
    g_inj.excl.push_back((int)v);
    p = (*end == ',') ? end + 1 : end;
  }
  for (int i = 0; i < nelements; ++i) {
    bool include = true;
    for (size_t j = 0; j < g_inj.excl.size(); ++j)
      if (pid[i] == g_inj.excl[j]) include = false;
    if (include) g_inj.elems.push_back(i);
  }
  const size_t n = g_inj.elems.size();
  g_inj.gt15.assign(n, 0); g_inj.gt30.assign(n, 0); g_inj.r120.assign(n, 0); g_inj.xsr28.assign(n, 0);
  g_inj.psOld.assign(n, 0.0); g_inj.psxsr.assign(n, 0.0);
}

static void InjuryStep() {
  const int n = (int)g_inj.elems.size();
  for (int j = 0; j < n; ++j) {
    const int i = g_inj.elems[j];
    double smax, smin, shear;
    CalculateMaximumPrincipalStrain(i, &smax, &smin, &shear); /* the reference's own */
    if (g_inj.maxStrain < smax) { g_inj.maxStrain = smax; g_inj.maxElem = i; g_inj.maxT = Time; }
    if (g_inj.minStrain > smin) { g_inj.minStrain = smin; g_inj.minElem = i; g_inj.minT = Time; }
    if (g_inj.maxShear < shear) { g_inj.maxShear = shear; g_inj.shearElem = i; g_inj.maxShearT = Time; }
    if (!g_inj.gt15[j] && smax > 0.15) g_inj.gt15[j] = 1;
    if (!g_inj.gt30[j] && smax > 0.30) g_inj.gt30[j] = 1;
    const double PSR = (smax - g_inj.psOld[j]) / dt;
    const double PSxSR = smax * PSR;
    if (g_inj.maxPSxSR < PSxSR) { g_inj.maxPSxSR = PSxSR; g_inj.maxElemPSxSR = i; g_inj.maxTimePSxSR = Time; }
    if (!g_inj.r120[j] && PSR > 120.0) g_inj.r120[j] = 1;
    if (!g_inj.xsr28[j] && PSxSR > 28.0) g_inj.xsr28[j] = 1;
    g_inj.psOld[j] = smax;
    g_inj.psxsr[j] = PSxSR;
  }
  const double v95 = compute95thPercentileValue(g_inj.psOld.data(), n); /* the reference's own (collective) */
  g_inj.hist95.push_back(v95);
  if (v95 > g_inj.max95) {
    g_inj.max95 = v95; g_inj.t95 = Time;
    g_inj.list95.clear();
    for (int j = 0; j < n; ++j) if (g_inj.psOld[j] >= g_inj.max95) g_inj.list95.push_back(g_inj.elems[j]);
  }
  const double x95 = compute95thPercentileValue(g_inj.psxsr.data(), n);
  g_inj.histx95.push_back(x95);
  if (x95 > g_inj.maxx95) {
    g_inj.maxx95 = x95; g_inj.tx95 = Time;
    g_inj.listx95.clear();
    for (int j = 0; j < n; ++j) if (g_inj.psxsr[j] >= g_inj.maxx95) g_inj.listx95.push_back(g_inj.elems[j]);
  }
}

static void InjuryDump() {
  const int n = (int)g_inj.elems.size();
  puti("inj_elems", g_inj.elems.data(), n);
  puti("inj_gt15", g_inj.gt15.data(), n); puti("inj_gt30", g_inj.gt30.data(), n);
  puti("inj_r120", g_inj.r120.data(), n); puti("inj_xsr28", g_inj.xsr28.data(), n);
  putd("inj_ps_old", g_inj.psOld.data(), n); putd("inj_psxsr", g_inj.psxsr.data(), n);
  puti("inj_list95", g_inj.list95.data(), (int64_t)g_inj.list95.size());
  puti("inj_listx95", g_inj.listx95.data(), (int64_t)g_inj.listx95.size());
  putd("inj_hist95", g_inj.hist95.data(), (int64_t)g_inj.hist95.size());
  putd("inj_histx95", g_inj.histx95.data(), (int64_t)g_inj.histx95.size());
  const double sc[12] = {g_inj.maxStrain, g_inj.maxT, g_inj.minStrain, g_inj.minT, g_inj.maxShear, g_inj.maxShearT,
                         g_inj.maxPSxSR, g_inj.maxTimePSxSR, g_inj.max95, g_inj.t95, g_inj.maxx95, g_inj.tx95};
  putd("inj_scalars", sc, 12);
  const int el[4] = {g_inj.maxElem, g_inj.minElem, g_inj.shearElem, g_inj.maxElemPSxSR};
  puti("inj_extreme_elems", el, 4);
  /* ex5.cpp:1043-1066: volumes by the reference's computePartVolume */
  std::vector<double> volumePart(nPIDglobal, 0.0), elementVolume(nelements > 0 ? nelements : 1, 0.0);
  computePartVolume(volumePart.data(), elementVolume.data());
  double vol[5] = {0, 0, 0, 0, 0};
  for (int j = 0; j < n; ++j) {
    const double eV = elementVolume[g_inj.elems[j]];
    if (g_inj.gt15[j]) { vol[0] += eV; if (g_inj.gt30[j]) vol[1] += eV; }
    if (g_inj.r120[j]) vol[2] += eV;
    if (g_inj.xsr28[j]) vol[3] += eV;
    vol[4] += eV;
  }
  putd("inj_volumes", vol, 5);
  putd("inj_volume_part", volumePart.data(), nPIDglobal);
}

/* ---- rigid-body prescribed motion of the brain drivers (ex5.cpp:339-371,574-912,976-1020): driver code in the
 * reference, restated here around the reference library's own quaternionExp / quaternionInverse / quaternionRotate /
 * crossProduct / dotProduct3D / interpolateLinear.  The driver's integrator is boost::numeric::odeint's
 * runge_kutta_dopri5 (Boost is neither in this image nor vendored): its do_step(sys, y, ydot, t, dt) is restated from
 * the published Dormand-Prince tableau -- the one piece of this path that is NOT the reference's own code. */
struct Rigid {
  bool on = false;
  std::vector<double> t[6], v[6];  // 0..2 angular x,y,z; 3..5 linear x,y,z (seconds, SI)
  double y[12], ydot[12];
  std::vector<int> boundaryID;
} g_rb;

static void rbDerivatives(const double *y, double *ydot, const double t) {  /* ex5.cpp:976-1020 */
  for (int k = 0; k < 3; ++k) {
    ydot[k] = interpolateLinear((int)g_rb.t[k].size(), g_rb.t[k].data(), g_rb.v[k].data(), t);
    ydot[6 + k] = interpolateLinear((int)g_rb.t[3 + k].size(), g_rb.t[3 + k].data(), g_rb.v[3 + k].data(), t);
    ydot[9 + k] = y[6 + k];
  }
  double r[3] = {y[3], y[4], y[5]};
  double *rdot = &ydot[3];
  double rMagnitude = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (rMagnitude < 1e-10) {
    rdot[0] = 0.5 * y[0]; rdot[1] = 0.5 * y[1]; rdot[2] = 0.5 * y[2];
  } else {
    double rCotR = rMagnitude / tan(rMagnitude);
    double omega[3] = {y[0], y[1], y[2]};
    crossProduct(omega, r, rdot);
    for (int i = 0; i < 3; ++i) r[i] = r[i] / rMagnitude;
    double rDotOmega = dotProduct3D(r, omega);
    for (int i = 0; i < 3; ++i) rdot[i] = 0.5 * (rdot[i] + rCotR * y[i] + (1.0 - rCotR) * rDotOmega * r[i]);
  }
}
static void rbDopri5(double *x, double *dxdt, double t, double dt) {
  static const double a2 = 1.0 / 5, a3 = 3.0 / 10, a4 = 4.0 / 5, a5 = 8.0 / 9, b21 = 1.0 / 5, b31 = 3.0 / 40, b32 = 9.0 / 40,
                      b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9, b51 = 19372.0 / 6561, b52 = -25360.0 / 2187,
                      b53 = 64448.0 / 6561, b54 = -212.0 / 729, b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247,
                      b64 = 49.0 / 176, b65 = -5103.0 / 18656, c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192,
                      c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  double xt[12], k2[12], k3[12], k4[12], k5[12], k6[12];
  const double *k1 = dxdt;
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b21 * k1[i];
  rbDerivatives(xt, k2, t + dt * a2);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b31 * k1[i] + dt * b32 * k2[i];
  rbDerivatives(xt, k3, t + dt * a3);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b41 * k1[i] + dt * b42 * k2[i] + dt * b43 * k3[i];
  rbDerivatives(xt, k4, t + dt * a4);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * b51 * k1[i] + dt * b52 * k2[i] + dt * b53 * k3[i] + dt * b54 * k4[i];
  rbDerivatives(xt, k5, t + dt * a5);
  for (int i = 0; i < 12; ++i)
    xt[i] = 1.0 * x[i] + dt * b61 * k1[i] + dt * b62 * k2[i] + dt * b63 * k3[i] + dt * b64 * k4[i] + dt * b65 * k5[i];
  rbDerivatives(xt, k6, t + dt);
  for (int i = 0; i < 12; ++i) xt[i] = 1.0 * x[i] + dt * c1 * k1[i] + dt * c3 * k3[i] + dt * c4 * k4[i] + dt * c5 * k5[i] + dt * c6 * k6[i];
  for (int i = 0; i < 12; ++i) x[i] = xt[i];
  rbDerivatives(x, dxdt, t + dt);
}
static void RigidInit(const char *file) {  /* tables: 6 lines "n t0 v0 t1 v1 ..."; ex5.cpp:819-911 for the node set */
  FILE *f = fopen(file, "r");
  if (!f) { fprintf(stderr, "cannot open %s\n", file); TerminateFemTech(3); }
  for (int k = 0; k < 6; ++k) {
    int n = 0;
    if (fscanf(f, "%d", &n) != 1) TerminateFemTech(3);
    g_rb.t[k].resize(n); g_rb.v[k].resize(n);
    for (int i = 0; i < n; ++i)
      if (fscanf(f, "%lf %lf", &g_rb.t[k][i], &g_rb.v[k][i]) != 2) TerminateFemTech(3);
  }
  fclose(f);
  g_rb.on = true;
  for (int i = 0; i < nelements; ++i)
    if (materialID[pid[i]] == 0)
      for (int j = eptr[i]; j < eptr[i + 1]; ++j)
        for (int c = 0; c < 3; ++c) boundary[connectivity[j] * ndim + c] = 1;
  for (int i = 0; i < nNodes; ++i)
    if (boundary[i * ndim]) g_rb.boundaryID.push_back(i);
  for (int j = 0; j < 12; ++j) { g_rb.y[j] = 0.0; g_rb.ydot[j] = 0.0; }
  for (size_t i = 0; i < g_rb.boundaryID.size(); ++i)
    for (int j = 0; j < ndim; ++j) {
      const int index = g_rb.boundaryID[i] * ndim + j;
      displacements[index] = 0.0; velocities[index] = 0.0; accelerations[index] = 0.0;
    }
}
static void ApplyAccBoundaryConditions() {  /* ex5.cpp:339-371 */
  double r[4], R[4], Rinv[4], V[4], Vp[4];
  double omegaR[3], omega[3], omegaOmegaR[3], omegaVel[3], vel[3], alpha[3], alphaR[3], locV[3];
  double *yInt = g_rb.y, *ydotInt = g_rb.ydot;
  rbDopri5(yInt, ydotInt, Time - dt, dt);
  r[0] = 0.0; r[1] = yInt[3]; r[2] = yInt[4]; r[3] = yInt[5];
  quaternionExp(r, R);
  quaternionInverse(R, Rinv);
  omega[0] = yInt[0]; omega[1] = yInt[1]; omega[2] = yInt[2];
  alpha[0] = ydotInt[0]; alpha[1] = ydotInt[1]; alpha[2] = ydotInt[2];
  vel[0] = yInt[6]; vel[1] = yInt[7]; vel[2] = yInt[8];
  for (size_t i = 0; i < g_rb.boundaryID.size(); i++) {
    int index = g_rb.boundaryID[i] * ndim;
    for (int j = 0; j < ndim; ++j) locV[j] = coordinates[index + j];
    V[0] = 0.0; V[1] = locV[0]; V[2] = locV[1]; V[3] = locV[2];
    quaternionRotate(V, R, Rinv, Vp);
    crossProduct(omega, &(Vp[1]), omegaR);
    crossProduct(omega, omegaR, omegaOmegaR);
    crossProduct(omega, vel, omegaVel);
    crossProduct(alpha, &(Vp[1]), alphaR);
    for (int j = 0; j < ndim; ++j) {
      displacements[index + j] = Vp[j + 1] - locV[j] + yInt[9 + j];
      velocities[index + j] = omegaR[j] + yInt[6 + j];
      accelerations[index + j] = 2.0 * omegaVel[j] + omegaOmegaR[j] + ydotInt[6 + j] + alphaR[j];
    }
  }
}

/* Benchmarking-Parallel.cpp:184-244, cube side parametrised */
static void ApplyBoundaryConditions(double dMax, double tMax) {
  double tol = 1e-5;
  double AppliedDisp = Time * (dMax / tMax);
  int index;
  for (int i = 0; i < nNodes; i++) {
    index = ndim * i + 0;
    if (fabs(coordinates[index] - 0.0) < tol) {
      boundary[index] = 1;
      displacements[index] = 0.0;
      velocities[index] = 0.0;
      accelerations[index] = 0.0;
    }
    index = ndim * i + 1;
    if (fabs(coordinates[index] - 0.0) < tol) {
      boundary[index] = 1;
      displacements[index] = 0.0;
      velocities[index] = 0.0;
      accelerations[index] = 0.0;
    }
    index = ndim * i + 2;
    if (fabs(coordinates[index] - 0.0) < tol) {
      boundary[index] = 1;
      displacements[index] = 0.0;
      velocities[index] = 0.0;
      accelerations[index] = 0.0;
    }
    index = ndim * i + 1;
    if (fabs(coordinates[index] - g_cubeL) < tol) {
      boundary[index] = 1;
      displacements[index] = AppliedDisp;
      velocities[index] = dMax / tMax;
      accelerations[index] = 0.0;
    }
  }
}

static double now() {
  struct timeval tv;
  gettimeofday(&tv, 0);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s mesh out-prefix maxSteps tMax dMax [cubeL]\n", argv[0]);
    return 2;
  }
  const char *outPrefix = argv[2];
  const int maxSteps = atoi(argv[3]);
  const double tMax = atof(argv[4]);
  const double dMax = atof(argv[5]);
  if (argc > 6) g_cubeL = atof(argv[6]);
  const int warmupSteps = argc > 7 ? atoi(argv[7]) : 0;
  const bool nodump = argc > 8 && strcmp(argv[8], "nodump") == 0;

  double t0 = now();
  InitFemTechWoInput(argc, argv);
  ReadInputFile(argv[1]);
  ReadMaterials();
  PartitionMesh();
  AllocateArrays();

  char fname[1024];
  snprintf(fname, sizeof fname, "%s.rank%d.bin", outPrefix, world_rank);
  g_out = fopen(fname, "wb");
  if (!g_out) { fprintf(stderr, "cannot open %s\n", fname); TerminateFemTech(3); }

  g_nodump = nodump;
  puti1("world_size", world_size);
  puti1("world_rank", world_rank);
  puti1("nNodes", nNodes);
  puti1("nelements", nelements);
  puti1("nallelements", nallelements);
  puti1("nPIDglobal", nPIDglobal);
  putd("coordinates", coordinates, (int64_t)ndim * nNodes);
  puti("connectivity", connectivity, eptr[nelements]);
  puti("eptr", eptr, nelements + 1);
  puti("pid", pid, nelements);
  puti("global_eid", global_eid, nelements);
  puti("globalNodeID", globalNodeID, nNodes);
  puti("materialID", materialID, nPIDglobal);
  putd("properties", properties, (int64_t)nPIDglobal * MAXMATPARAMS);
  puti1("sendProcessCount", sendProcessCount);
  puti("sendProcessID", sendProcessID, sendProcessCount);
  puti("sendNeighbourCount", sendNeighbourCount, sendProcessCount);
  puti("sendNeighbourCountCum", sendNeighbourCountCum, sendProcessCount + 1);
  puti("sendNodeIndex", sendNodeIndex, sendProcessCount ? sendNeighbourCountCum[sendProcessCount] : 0);

  Time = 0.0;
  dt = 0.0;
  ShapeFunctions();
  putd("shp", shp, gptr[nelements]);
  putd("dshp", dshp, dsptr[nelements]);
  putd("detJacobian", detJacobian, gpPtr[nelements]);
  putd("gaussWeights", gaussWeights, gpPtr[nelements]);
  AssembleLumpedMass();
  putd("mass", mass, nDOF);

  if (const char *rf = getenv("REF_RIGID")) RigidInit(rf);  /* ex5.cpp:136 InitBoundaryCondition */
  else ApplyBoundaryConditions(dMax, tMax);
  puti("boundary0", boundary, nDOF);
  dt = ExplicitTimeStepReduction * StableTimeStep();
  GetForce();
  CalculateAccelerations();
  puts1("dt0", dt);
  putd("fi0", fi, nDOF);
  putd("accelerations0", accelerations, nDOF);
  putd("pk2_0", pk2, pk2ptr[nelements]);

  if (const char *spec = getenv("REF_INJURY_EXCLUDE")) InjuryInit(spec);
  std::vector<double> dtHist;
  int time_step_counter = 1;
  double t_n = 0.0;
  double tSetup = now() - t0;
  double tLoop0 = now();
  int steps = 0;
  while (Time < tMax && steps < maxSteps) {
    if (steps == warmupSteps) tLoop0 = now();
    t_n = Time;
    double t_np1 = Time + dt;
    Time = t_np1;
    double dt_nphalf = dt;
    double t_nphalf = 0.5 * (t_np1 + t_n);
    dtHist.push_back(dt);
    for (int i = 0; i < nDOF; i++) {
      if (boundary[i]) {
        velocities_half[i] = velocities[i];
      } else {
        velocities_half[i] = velocities[i] + (t_nphalf - t_n) * accelerations[i];
      }
    }
    memcpy(displacements_prev, displacements, nDOF * sizeof(double));
    memcpy(accelerations_prev, accelerations, nDOF * sizeof(double));
    memcpy(fi_prev, fi, nDOF * sizeof(double));
    memcpy(fe_prev, fe, nDOF * sizeof(double));
    for (int i = 0; i < nDOF; i++) {
      if (!boundary[i]) {
        displacements[i] = displacements[i] + dt_nphalf * velocities_half[i];
      }
    }
    if (g_rb.on) ApplyAccBoundaryConditions();  /* ex5.cpp:222 */
    else ApplyBoundaryConditions(dMax, tMax);
    GetForce();
    CalculateAccelerations();
    for (int i = 0; i < nDOF; i++) {
      if (!boundary[i]) {
        velocities[i] = velocities_half[i] + (t_np1 - t_nphalf) * accelerations[i];
      }
    }
    CheckEnergy(Time, 0); /* writeFlag 0: one energy_<uid>.dat line per step */
    if (g_inj.on) InjuryStep();  /* ex5.cpp:240 */
    time_step_counter = time_step_counter + 1;
    steps++;
    dt = ExplicitTimeStepReduction * StableTimeStep();
    MPI_Barrier(MPI_COMM_WORLD);
  }
  double tLoop = now() - tLoop0;
  CalculateStrain();

  puti1("steps", steps);
  puts1("Time", Time);
  puts1("dt", dt);
  putd("dt_hist", dtHist.data(), (int64_t)dtHist.size());
  putd("displacements", displacements, nDOF);
  putd("velocities", velocities, nDOF);
  putd("velocities_half", velocities_half, nDOF);
  putd("accelerations", accelerations, nDOF);
  puti("boundary", boundary, nDOF);
  putd("fi", fi, nDOF);
  putd("f_net", f_net, nDOF);
  putd("F", F, fptr[nelements]);
  putd("detF", detF, detFptr[nelements]);
  putd("pk2", pk2, pk2ptr[nelements]);
  putd("Eavg", Eavg, (int64_t)nelements * 9);
  putd("Hn_1", Hn_1, Hn_1 ? fptr[nelements] : 0);
  putd("Hn_2", Hn_2, Hn_2 ? fptr[nelements] : 0);
  putd("S0n", S0n, S0n ? fptr[nelements] : 0);
  if (g_inj.on) InjuryDump();
  if (g_rb.on) {
    putd("rb_y", g_rb.y, 12); putd("rb_ydot", g_rb.ydot, 12);
    puti("rb_boundaryID", g_rb.boundaryID.data(), (int64_t)g_rb.boundaryID.size());
  }
  puts1("wall_setup_s", tSetup);
  puts1("wall_loop_s", tLoop);
  fclose(g_out);
  if (world_rank == 0) {
    const int timed = steps > warmupSteps ? steps - warmupSteps : 0;
    printf("REF_DUMP ranks=%d nallelements=%d steps=%d timed_steps=%d Time=%.17g dt=%.17g setup_s=%.3f loop_s=%.6f "
           "element_steps_per_s=%.6e u0=(%.9e %.9e %.9e) uid=%s\n",
           world_size, nallelements, steps, timed, Time, dt, tSetup, tLoop,
           tLoop > 0 ? (double)nallelements * timed / tLoop : 0.0, displacements[0], displacements[1],
           displacements[2], uid.c_str());
    fflush(stdout);
  }
  FinalizeFemTech();
  return 0;
}
