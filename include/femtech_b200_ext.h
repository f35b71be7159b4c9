/*
 * femtech_b200_ext.h -- the ONE addition to the reference's API that the resident mode needs.
 *
 * The reference's ExplicitDynamics(double timeFinal, char *name) (include/FemTech.h:50) carries no boundary
 * condition: the shipped drivers apply it from a callback inside their own time loop
 * (examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:137,184-244).  With the loop resident on the GPU the
 * callback becomes a descriptor, handed over once before ExplicitDynamics():
 *   bc_kind[3*nNodes]  0 = free dof, k in 1..3 = prescribed: u = Time*bc_rate[k], v = bc_rate[k], a = 0, boundary = 1
 *   energy_every       1 = run the energy check every step like the drivers do, 0 = never
 */
#ifndef FEMTECH_B200_EXT_H
#define FEMTECH_B200_EXT_H
void femtech_b200_set_bc(const int *bc_kind, const double bc_rate[4], int energy_every);
#endif
