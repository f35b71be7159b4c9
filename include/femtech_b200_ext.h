/*
 * femtech_b200_ext.h -- the additions to the reference's API that the resident mode needs.
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

/* The brain drivers evaluate their injury criteria from a second callback in the same loop
 * (examples/ex5/ex5.cpp:240 CalculateInjuryCriterions, :1251 InitInjuryCriterion; driver code, not library code).
 * Resident mode: call femtech_b200_injury_begin() once after ShapeFunctions(); every step of ExplicitDynamics() then
 * updates the criteria on the device; read them back with femtech_b200_injury_results() (arguments as
 * ftb200_injury_get in ftb200.h: the driver's maxStrain/maxT/..., MPSgt15/.. as flag bits, PS_Old, PSxSRArray,
 * flagged volumes).  Strict legacy drivers need nothing new: CalculateMaximumPrincipalStrain(e, ...) keeps its
 * signature (include/FemTech.h:60) and is served from one device evaluation per force call. */
void femtech_b200_injury_begin(const int *injuryExcludePID, int injuryExcludePIDCount);

/* The brain drivers' boundary condition is a third callback, ApplyAccBoundaryConditions (ex5.cpp:222,339-371): the
 * nodes of the rigid part follow a prescribed rigid-body motion integrated from acceleration traces.  Resident mode:
 * hand the traces over once (arguments as ftb200_set_rigid_bc in ftb200.h: angular x,y,z then linear x,y,z traces in
 * s and SI units as InitBoundaryCondition leaves them, ex5.cpp:577-645; boundaryID NULL = nodes of material-0 parts). */
void femtech_b200_set_rigid_bc(const int sizes[6], const double *const t[6], const double *const v[6], const int *boundaryID,
                               int boundarySize, int energy_every);
void femtech_b200_injury_results(double scalars12[12], int extreme_elems4[4], unsigned char *flags, double *PS_Old,
                                 double *PSxSRArray, double volumes5[5]);
/* Steps executed and seconds spent inside the last ExplicitDynamics() call (set-up of step 0 and the final read-back
 * included), the transport used on several ranks ("single", "p2p", "host"): what a driver prints as its throughput. */
long long femtech_b200_last_steps(void);
double femtech_b200_last_seconds(void);
const char *femtech_b200_last_transport(void);
#endif
