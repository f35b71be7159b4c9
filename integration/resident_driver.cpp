// resident_driver.cpp -- a FemTech driver in the style of examples/Benchmarking-Parallel that uses the RESIDENT mode of
// femtech_b200: the reference's own setup calls (InitFemTechWoInput, ReadInputFile, ReadMaterials, PartitionMesh,
// AllocateArrays, ShapeFunctions, AssembleLumpedMass: include/FemTech.h), the boundary condition of the benchmark
// (Benchmarking-Parallel.cpp:184-244) handed over as a descriptor, then ONE ExplicitDynamics(tMax) -- the whole time
// loop of Benchmarking-Parallel.cpp:106-171 runs on the GPU and the energy file receives one line per step, written from
// the records the device streams into pinned host memory.  Linked by integration/build_dropin.sh against
// femtech_host.o + libftb200.so + the reference's readers/partitioner.
//
//   resident_driver <mesh.inp|.k> [tMax=0.1] [dMax=0.007] [cubeL=0.005] [energy_every=1]
//
// Writes <mesh>.resident.txt: nNodes, steps (from the energy file), Time, then u of every node (%.17g) for the test.
#include "FemTech.h"
#include "femtech_b200_ext.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

double Time, dt;
int nSteps;
double ExplicitTimeStepReduction = 0.8;
double FailureTimeStep = 1e-11;
int nPlotSteps = 50;
bool ImplicitStatic = false;
bool ImplicitDynamic = false;
bool ExplicitDynamic = true;

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s mesh [tMax] [dMax] [cubeL] [energy_every]\n", argv[0]);
    return 2;
  }
  const double tMax = argc > 2 ? atof(argv[2]) : 0.1, dMax = argc > 3 ? atof(argv[3]) : 0.007;
  const double L = argc > 4 ? atof(argv[4]) : 0.005;
  const int energy_every = argc > 5 ? atoi(argv[5]) : 1;
  InitFemTechWoInput(argc, argv);
  ReadInputFile(argv[1]);
  ReadMaterials();
  PartitionMesh();
  AllocateArrays();
  Time = 0.0;
  dt = 0.0;
  ShapeFunctions();
  AssembleLumpedMass();
  // the benchmark's boundary condition as a descriptor: faces x = 0, y = 0, z = 0 held in their normal direction,
  // face y = L pulled with u_y = Time * dMax / tMax
  const double tol = 1e-5;
  std::vector<int> kind((size_t)nNodes * ndim, 0);
  double rate[4] = {0.0, 0.0, dMax / tMax, 0.0};
  for (int i = 0; i < nNodes; ++i) {
    for (int c = 0; c < ndim; ++c)
      if (fabs(coordinates[ndim * i + c]) < tol) kind[ndim * i + c] = 1;
    if (fabs(coordinates[ndim * i + 1] - L) < tol) kind[ndim * i + 1] = 2;
  }
  femtech_b200_set_bc(kind.data(), rate, energy_every);
  ExplicitDynamics(tMax, argv[1]);
  // one file per rank when there are several (with the global node ids of PartitionMesh.cpp, so that a test can place them)
  std::string out = std::string(argv[1]) + ".resident.txt";
  if (world_size > 1) out = std::string(argv[1]) + ".resident.rank" + std::to_string(world_rank) + ".txt";
  FILE *f = fopen(out.c_str(), "w");
  if (!f) return 3;
  fprintf(f, "%d %.17g %.17g\n", nNodes, Time, dt);
  for (int i = 0; i < nNodes * ndim; ++i) fprintf(f, "%.17g\n", displacements[i]);
  if (world_size > 1) {
    for (int i = 0; i < nNodes; ++i) fprintf(f, "%d\n", globalNodeID[i]);
    for (int i = 0; i < nNodes * ndim; ++i) fprintf(f, "%.17g\n", velocities[i]);
  }
  fclose(f);
  {  // throughput of the call, summed over the ranks' elements (every element belongs to exactly one rank)
    long long eLocal = nelements, eTotal = 0;
    MPI_Reduce(&eLocal, &eTotal, 1, MPI_LONG_LONG, MPI_SUM, 0, MPI_COMM_WORLD);
    double sec = femtech_b200_last_seconds(), secMax = 0.0;
    MPI_Reduce(&sec, &secMax, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
    if (world_rank == 0)
      printf("RESIDENT ranks %d transport %s elements %lld steps %lld seconds %.6f element_steps_per_s %.6e\n", world_size,
             femtech_b200_last_transport(), eTotal, femtech_b200_last_steps(), secMax,
             (double)eTotal * (double)femtech_b200_last_steps() / secMax);
  }
  FinalizeFemTech();  // frees the arrays (InitFinalizeFemTech.cpp:75-80)
  return 0;
}
