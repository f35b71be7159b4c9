/*
 * ftb200.h -- C-ABI of the B200-native FemTech explicit-dynamics hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, raw pointers and sizes,
 * int status returns, no C++/torch types.  The C++ host layer that mirrors the
 * reference's API (femtech_b200/csrc/femtech_host.cpp: ShapeFunctions(),
 * AssembleLumpedMass(), GetForce(), CalculateAccelerations(), StableTimeStep(),
 * CheckEnergy(), ExplicitDynamics() over the GlobalVariables.h arrays) and the
 * Python mirror used by tests/bench (femtech_b200/solver.py) both sit on top of
 * exactly these entry points.  Each one cites the reference interface it
 * replaces (paths relative to the FemTech repository root).
 *
 * Conventions shared with the reference (include/GlobalVariables.h:18-127):
 *   nodal arrays are AoS, xyz interleaved, double[3*nNodes]; boundary is
 *   int[3*nNodes]; connectivity is int[8*nElements] of LOCAL node ids in C3D8
 *   order; properties is double[9*nPID]; PK2 stress is Voigt [11,22,33,23,13,12]
 *   per Gauss point, F is column-major 3x3 per Gauss point.
 * Host pointers may be pageable or pinned.  All calls are synchronous with
 * respect to the host unless stated otherwise, single-threaded per context.
 *
 * Return value: 0 on success, otherwise the reference's TerminateFemTech code
 * for the same failure (src/io/InitFinalizeFemTech.cpp:82-86): 1 unknown
 * material, 3 bad input, 12 allocation failure, 19 time step below
 * FailureTimeStep; or FTB200_ERR_CUDA for a CUDA runtime error.  The message is
 * available from ftb200_last_error().  There is no CPU fallback: without a
 * usable CUDA device ftb200_create() fails.
 */
#ifndef FTB200_H
#define FTB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FTB200_OK 0
#define FTB200_ERR_MATERIAL 1
#define FTB200_ERR_INPUT 3
#define FTB200_ERR_ALLOC 12
#define FTB200_ERR_TIMESTEP 19
#define FTB200_ERR_CUDA 100

typedef struct ftb200_ctx ftb200_ctx;

/* ---- lifetime ----------------------------------------------------------- */
/* One context per rank/GPU (reference: one MPI rank, InitFinalizeFemTech.cpp:49-73). */
int ftb200_create(int rank, int nranks, int device, ftb200_ctx **out);
int ftb200_destroy(ftb200_ctx *ctx);
const char *ftb200_last_error(const ftb200_ctx *ctx);
/* Run all kernels on the caller's CUDA stream (cudaStream_t as void*); NULL = own stream. */
int ftb200_set_stream(ftb200_ctx *ctx, void *cuda_stream);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long ftb200_launch_count(const ftb200_ctx *ctx);
/* Library build info: "sm_100a ..." */
const char *ftb200_build_info(void);

/* ---- setup (outputs of ReadInputFile / PartitionMesh / ReadMaterials) ---- */
/* coordinates, connectivity, pid: GlobalVariables.h:32-33,47-49 after PartitionMesh.cpp:485-536 */
int ftb200_upload_mesh(ftb200_ctx *ctx, const double *coordinates, const int *connectivity, const int *pid,
                       int nNodes, int nElements);
/* Mixed meshes of the reference's two solid elements: connectivity packed 8 (C3D8) or 4 (C3D4)
 * node ids per element as in include/GlobalVariables.h:32-33 with eptr[nElements+1] (ReadAbaqus.cpp:113-146).  C3D4:
 * one Gauss point, src/fem/ShapeFunctions/ShapeFunction_C3D4.cpp, CalculateCharacteristicLength_C3D4.cpp.  The lazy
 * outputs then use the reference's packed layouts: F[9*nGP], detF[nGP], pk2[6*nGP], nGP = ftb200_gauss_point_count. */
int ftb200_upload_mesh_mixed(ftb200_ctx *ctx, const double *coordinates, const int *connectivity, const int *eptr,
                             const int *pid, int nNodes, int nElements);
long long ftb200_gauss_point_count(ftb200_ctx *ctx);
/* Hexahedra of the mesh whose reference geometry is a parallelepiped (valid after ftb200_shape_functions; -1 before).
 * They are integrated by the kernel that forms dN/dX once per element instead of once per Gauss point
 * (ShapeFunction_C3D8.cpp:60-115 stores it per point); FTB200_AFFINE=0 in the environment sends them through the
 * general kernel.  Reported by bench.py next to the throughput. */
long long ftb200_affine_element_count(ftb200_ctx *ctx);
/* Brick decomposition of the brick-fused step (valid after ftb200_shape_functions): out8 = bricks, interior nodes
 * (finished inside the brick's thread block), surface nodes (finished by the second, thin pass), partial-sum slots,
 * brick dimensions in elements (3), 1 if the resident loop will take the brick-fused step.  bricks = 0: the mesh does not
 * qualify (several partitions, tetrahedra, distorted hexahedra, materials other than 1 and 4) or FTB200_BRICK=0, and
 * the two-kernel step (element forces through HBM, GetForce_3D.cpp:15-51 as two passes) is used. */
int ftb200_brick_info(ftb200_ctx *ctx, long long *out8);
/* The decomposition itself, for inspection and tests: brick_of_element[nElements] and, per node, the brick that finishes
 * it or -1 for a surface node (caller's numbering).  Either pointer may be NULL. */
int ftb200_brick_maps(ftb200_ctx *ctx, int *brick_of_element, int *interior_brick_of_node);
/* materialID, properties: src/io/input/ReadMaterials.cpp:8-138 */
int ftb200_upload_materials(ftb200_ctx *ctx, const int *materialID, const double *properties, int nPID);
/* sendProcessID / sendNeighbourCountCum / sendNodeIndex: PartitionMesh.cpp:566-1128 */
int ftb200_upload_comm(ftb200_ctx *ctx, int sendProcessCount, const int *sendProcessID,
                       const int *sendNeighbourCountCum, const int *sendNodeIndex);
/* ShapeFunctions() (src/fem/ShapeFunctions/ShapeFunctions.cpp:32-255): validates the mesh (positive
 * reference Jacobians), builds the node->element CSR map, zeroes the Prony history.  Nothing per
 * Gauss point is stored.  min_detJ (optional) receives the smallest reference detJ. */
int ftb200_shape_functions(ftb200_ctx *ctx, double *min_detJ);
/* AssembleLumpedMass() without the neighbour sum (src/fem/Mass/Mass3D.cpp:127-157).  mass_out (optional)
 * is double[3*nNodes].  Multi-GPU: sum shared nodes with the halo calls below (field 1). */
int ftb200_lumped_mass(ftb200_ctx *ctx, double *mass_out);
/* Current nodal mass (after any neighbour sum), double[3*nNodes] like the reference's `mass` array. */
int ftb200_get_mass(ftb200_ctx *ctx, double *mass_out);

/* ---- legacy per-call path: host arrays in, host arrays out --------------- */
/* GetForce()/GetForce_3D() (src/fem/SolidMechanics/GetForce_3D.cpp:5-53).  dt is the driver global `dt`
 * (only material 5 reads it).  fe may be NULL (== 0).  With a comm pattern uploaded and nranks > 1 this
 * call computes the LOCAL part only; use ftb200_halo_pack / ftb200_halo_add (field 0) for the neighbour sum. */
int ftb200_get_force(ftb200_ctx *ctx, const double *displacements, const double *fe, double dt, double *fi,
                     double *f_net);
/* CalculateAccelerations() (src/fem/SolidMechanics/CalculateAcclerations.cpp:4-13): accelerations[i] =
 * f_net[i]/mass[i] where !boundary[i]; other entries of the host array are left untouched. */
int ftb200_calculate_accelerations(ftb200_ctx *ctx, const int *boundary, double *accelerations);
/* StableTimeStep() local part (src/timestep/StableTimeStep.cpp:11-30): min over elements with a node that
 * is not fully constrained.  The caller applies the cross-rank MIN and the FailureTimeStep test. */
int ftb200_stable_time_step(ftb200_ctx *ctx, const double *displacements, const int *boundary, double *dtMin);
/* CheckEnergy() local sums (src/fem/SolidMechanics/CheckEnergy.cpp:19-52): out[0..2] = WKE, Wint, Wext
 * increments (already x0.5) over the nodes this rank owns. */
int ftb200_check_energy(ftb200_ctx *ctx, const double *displacements, const double *displacements_prev,
                        const double *velocities, const double *accelerations, const double *accelerations_prev,
                        const double *fi, const double *fi_prev, const double *fe, const double *fe_prev,
                        const int *boundary, double out[3]);
/* Lazy outputs of the last force evaluation in the reference's layouts: F[72*nE], detF[8*nE], pk2[48*nE]
 * (GetForce_3D.cpp side effects), Eavg[9*nE] (CalculateStrain.cpp:77-97).  Any pointer may be NULL. */
int ftb200_get_gp_outputs(ftb200_ctx *ctx, double *F, double *detF, double *pk2, double *Eavg);

/* ---- resident (fused) path ------------------------------------------------ */
/* Upload / download nodal state; any pointer may be NULL (skipped). */
int ftb200_set_state(ftb200_ctx *ctx, const double *displacements, const double *velocities,
                     const double *accelerations, const int *boundary);
int ftb200_get_state(ftb200_ctx *ctx, double *displacements, double *velocities, double *accelerations,
                     int *boundary, double *fi, double *f_net);
/* Boundary-condition descriptor replacing the drivers' ApplyBoundaryConditions callback
 * (examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:184-244): bc_kind[3*nNodes] in 0..3;
 * kind k > 0 prescribes u = Time*bc_rate[k], v = bc_rate[k], a = 0 and sets boundary = 1. */
int ftb200_set_bc(ftb200_ctx *ctx, const int *bc_kind, const double bc_rate[4]);
/* Step 0 of the explicit drivers (Benchmarking-Parallel.cpp:83-91): apply BC at Time0, dt =
 * reduction*StableTimeStep(), GetForce(), CalculateAccelerations().  energy_every: 0 = never call the
 * energy check, k = every k-th step (the shipped drivers use 1). */
int ftb200_explicit_begin(ftb200_ctx *ctx, double Time0, double ExplicitTimeStepReduction, double FailureTimeStep,
                          int energy_every);
/* The time loop (Benchmarking-Parallel.cpp:106-171) on the device: runs while Time < tMax for at most
 * maxSteps steps.  This is ExplicitDynamics(timeFinal, name) (include/FemTech.h:50, a stub in the
 * reference).  Outputs (optional): steps executed, Time, next dt.  Asynchronous variant: _async enqueues
 * `steps` steps on the stream without any host synchronisation (bench timing); _poll reads the scalars. */
int ftb200_explicit_run(ftb200_ctx *ctx, double tMax, long long maxSteps, long long *steps_done, double *Time,
                        double *dt);
int ftb200_explicit_run_async(ftb200_ctx *ctx, double tMax, long long steps);
int ftb200_explicit_poll(ftb200_ctx *ctx, long long *steps_done, double *Time, double *dt, int *status_bits);
/* Non-blocking variant: enqueues a device->host copy of the step scalars behind the work already queued; out8 (pinned
 * host memory, 8 doubles) receives Time, dt, steps done, status bits, Wint, Wext, WKE, |balance| once the stream gets
 * there.  No host synchronisation. */
int ftb200_explicit_poll_async(ftb200_ctx *ctx, double *out8_pinned);
/* Step ring: the device itself writes one 8-double record per finished step -- Time, next dt, finished steps, status
 * bits, Wint, Wext, WKE, |balance| (the drivers' per-step log line, Benchmarking-Parallel.cpp:160-166 / CheckEnergy.cpp:66-83)
 * -- into pinned host memory owned by the library (*host_ring, capacity records, slot = (step - 1) % capacity), 64 bytes
 * over PCIe per step and no host synchronisation.  record[2] (the step counter) is stored last: a host that reads
 * record[2] == k may use the rest of step k's record.  capacity 0 releases the ring.  Standard loop (element + node
 * kernels per step, one or several partitions); the opt-in one-kernel variants do not write it. */
int ftb200_step_ring(ftb200_ctx *ctx, long long capacity, double **host_ring);
/* Running energies of the last checked step: out[0..3] = Wint, Wext, WKE, |WKE+Wint-Wext| */
int ftb200_get_energy(ftb200_ctx *ctx, double out[4]);
/* Optional per-step records kept on the device: capacity in steps (0 disables). */
int ftb200_record_history(ftb200_ctx *ctx, long long capacity);
int ftb200_get_history(ftb200_ctx *ctx, long long first, long long count, double *dt_hist, double *energy_hist4);

/* ---- multi-GPU: shared-node exchange (GetForce_3D.cpp:54-102, Mass3D.cpp:77-125) ---- */
/* Total number of shared-node slots (sendNeighbourCountCum[sendProcessCount]). */
int ftb200_halo_count(const ftb200_ctx *ctx);
/* Transport-agnostic split form.  Buffers are DEVICE pointers of 3*halo_count doubles laid out exactly
 * like the reference's sendNodeDisplacement / recvNodeDisplacement (neighbour-major, xyz interleaved), so
 * any transport (NCCL send/recv per neighbour, peer stores) can move slice i of send to the neighbour's
 * slice of recv.  field: 0 = internal force, 1 = lumped mass. */
int ftb200_halo_pack(ftb200_ctx *ctx, int field, double *send_dev);
int ftb200_halo_add(ftb200_ctx *ctx, int field, const double *recv_dev);
/* HOST-buffer variants for a host that moves the windows itself and has no CUDA of its own -- the reference's MPI ranks:
 * send_host / recv_host are the reference's sendNodeDisplacement / recvNodeDisplacement arrays (3*halo_count doubles),
 * exchanged by the caller with the very MPI_Isend / MPI_Irecv loop of GetForce_3D.cpp:63-91 / Mass3D.cpp:86-114.
 * The device-side windows live inside the context.  (integration/femtech_host.cpp under mpirun -np N.) */
int ftb200_device_count(void);
int ftb200_halo_pack_host(ftb200_ctx *ctx, int field, double *send_host);
int ftb200_halo_add_host(ftb200_ctx *ctx, int field, const double *recv_host);
/* GetForce_3D.cpp:5-102 on a rank with neighbours, around the exchange: element loop + local scatter + the partial sums
 * of the shared nodes packed (:15-61), then -- after the caller's exchange -- neighbour sum in ascending neighbour
 * order and f_net (:92-97, :49-51). */
int ftb200_get_force_begin(ftb200_ctx *ctx, const double *displacements, const double *fe, double dt, double *send_host);
int ftb200_get_force_end(ftb200_ctx *ctx, const double *recv_host, double *fi, double *f_net);
/* StableTimeStep.cpp:33: this rank's candidate out, MPI_Allreduce(MIN) by the caller, the global value back in. */
int ftb200_get_dtmin(ftb200_ctx *ctx, double *dtmin_local);
int ftb200_set_dtmin(ftb200_ctx *ctx, double dtmin_global);
/* explicit_begin and the time step of the resident loop with the exchange on the host (the sequence of
 * ftb200_explicit_begin_dt/_force/_finish and ftb200_step_begin/_join/_end below, host buffers instead of device ones):
 *   ftb200_explicit_begin_dt(.., NULL); ftb200_get_dtmin; [Allreduce MIN]; ftb200_set_dtmin;
 *   ftb200_explicit_begin_force_host(send); [exchange]; ftb200_explicit_begin_finish_host(recv);
 *   ftb200_run_begin(tMax, steps); per step: ftb200_step_begin_host(send, &dt_local); [exchange, Allreduce MIN];
 *   ftb200_step_end_host(recv, dt_global). */
int ftb200_explicit_begin_force_host(ftb200_ctx *ctx, double *send_host);
int ftb200_explicit_begin_finish_host(ftb200_ctx *ctx, const double *recv_host);
int ftb200_step_begin_host(ftb200_ctx *ctx, double *send_host, double *dtmin_local);
int ftb200_step_end_host(ftb200_ctx *ctx, const double *recv_host, double dtmin_global);
/* Resident path for a rank with shared nodes, split around the exchange (one call sequence per time step):
 *   ftb200_run_begin(tMax, steps)   once per run: arms the loop and performs the first kick + drift + BC
 *   ftb200_step_begin(send_dev, &dtmin_dev)
 *        element kernel on the elements that touch a shared node, partial f_int of the shared nodes packed
 *        into send_dev; the interior elements are launched on a second stream and overlap the exchange
 *   ... caller moves send_dev -> neighbours' recv_dev (any transport; same stream) ...
 *   ftb200_step_join()              the interior elements have been enqueued before what follows
 *   ... caller MIN-reduces *dtmin_dev across ranks in place (StableTimeStep.cpp:33) ...
 *   ftb200_step_end(recv_dev)       scalar update, then the node kernel adds recv_dev in ascending
 *                                   neighbour order (GetForce_3D.cpp:92-97) and finishes the step
 * dtmin_dev points at a device double.  Nothing here synchronises with the host. */
int ftb200_run_begin(ftb200_ctx *ctx, double tMax, long long steps);
/* explicit_begin for a rank with shared nodes, in three phases: _dt (BC + local element dt; caller then
 * MIN-reduces *dtmin_dev), _force (dt := reduction*min, GetForce, shared-node partials packed into send_dev;
 * caller exchanges), _finish (neighbour sum + CalculateAccelerations; synchronises and reports errors). */
int ftb200_explicit_begin_dt(ftb200_ctx *ctx, double Time0, double ExplicitTimeStepReduction, double FailureTimeStep,
                             int energy_every, double **dtmin_dev);
int ftb200_explicit_begin_force(ftb200_ctx *ctx, double *send_dev);
int ftb200_explicit_begin_finish(ftb200_ctx *ctx, const double *recv_dev);
int ftb200_step_begin(ftb200_ctx *ctx, double *send_dev, double **dtmin_dev);
int ftb200_step_join(ftb200_ctx *ctx);
int ftb200_step_end(ftb200_ctx *ctx, const double *recv_dev);
/* Peer-memory transport (NVLink/NVSwitch, no NCCL on the data path): every rank exports a device window,
 * imports the other ranks', and from then on ftb200_explicit_run* runs the whole step on a rank with shared
 * nodes -- pack kernel storing the shared-node partials straight into the neighbours' windows, flag-based
 * arrival, dt MIN through the same windows -- with no host or NCCL involvement (CUDA-graph captured).
 *   export: allocates the window; handle_out (64 bytes, may be NULL) receives the cudaIpcMemHandle_t,
 *           *window_out (may be NULL) the device pointer (for ranks living in the same process).
 *   import: all_handles = nranks entries in rank order, either 64-byte IPC handles (handles_are_pointers = 0)
 *           or device pointers (void*, handles_are_pointers = 1).  For every neighbour i of this rank
 *           (sendProcessID order): peer_slot_offset[i] = first slot of this rank's slice in that neighbour's
 *           receive window (its sendNeighbourCountCum entry for us), peer_my_index[i] = our index in its
 *           neighbour list, peer_halo_count[i] = its total slot count.  The send lists are symmetric
 *           (PartitionMesh.cpp:1071-1108), so these come from an all-gather of the comm patterns. */
#define FTB200_IPC_HANDLE_BYTES 64
int ftb200_p2p_export(ftb200_ctx *ctx, void *handle_out, void **window_out);
int ftb200_p2p_import(ftb200_ctx *ctx, const void *all_handles, int handles_are_pointers,
                      const int *peer_slot_offset, const int *peer_my_index, const int *peer_halo_count);

/* ---- measurement helpers --------------------------------------------------- */
/* Average device time per launch (ms) of the element and node kernels over the last explicit_run*,
 * measured with CUDA events on the launching stream when profiling is enabled (adds two events per
 * kernel; off by default). */
int ftb200_profile_enable(ftb200_ctx *ctx, int on);
int ftb200_profile_get(ftb200_ctx *ctx, double *elem_ms, double *node_ms, long long *elem_launches,
                       long long *node_launches);
/* Roofline denominators measured on this device: dependent-free DFMA chains on every SM (TFLOP/s, FMA = 2)
 * and a STREAM-style fp64 copy (GB/s, read + write bytes).  Best of `reps` launches, CUDA-event timed. */
int ftb200_measure_peaks(ftb200_ctx *ctx, int reps, double *fp64_tflops, double *copy_gbs);

/* ---- injury criteria of the brain drivers (next row after the hot path: SURVEY.md 8(f).1) ---------------
 * Device-resident restatement of examples/ex5/ex5.cpp:1251-1430 (InitInjuryCriterion, CalculateInjuryCriterions)
 * over src/elements/ElementCalculations/CalculateStrain.cpp:8-75 (CalculateMaximumPrincipalStrain) and
 * src/math/math.cpp:160-332 (compute95thPercentileValue).  Once begun, every step of the resident loop evaluates
 * the criteria inside the element force kernel (F never leaves the chip) right after the step's energy check,
 * as ex5.cpp:237-240 does.
 * exclude_pids: parts left out (ex5's injuryExcludePID).  thresholds: NULL = the reference's 0.15, 0.30 (MPS),
 * 120 1/s (MPSR), 28 1/s (MPSxSR), ex5.cpp:1335-1365.
 * Several partitions (nranks > 1): the flags, extrema and lists are per rank as in the reference, but the percentile is
 * a global order statistic (math.cpp:160-199 gathers every rank's array).  After injury_begin tell every rank the
 * global number of participating elements (sum of ftb200_injury_local_count), and after each ftb200_step_end run the
 * ftb200_injury_passes() radix passes: ftb200_injury_select_hist (local histogram into device memory, 2 x 2048
 * unsigned) -> sum over the ranks (e.g. NCCL all-reduce on *hist_dev) -> ftb200_injury_select_pick. */
int ftb200_injury_begin(ftb200_ctx *ctx, const int *exclude_pids, int n_exclude, const double *thresholds4);
int ftb200_injury_end(ftb200_ctx *ctx);
int ftb200_injury_local_count(ftb200_ctx *ctx, long long *n_included);
int ftb200_injury_global_count(ftb200_ctx *ctx, long long n_total);
int ftb200_injury_passes(void);
int ftb200_injury_select_hist(ftb200_ctx *ctx, int pass, unsigned **hist_dev, int *hist_len);
int ftb200_injury_select_pick(ftb200_ctx *ctx, int pass);
/* Results so far.  scalars[12] = maxStrain, time, minStrain, time, maxShear, time, maxPSxSR, time, MPS-95, time,
 * MPSxSR-95, time (ex5.cpp:62-83); extreme_elems[4] = the elements of the first four (caller's element ids);
 * flags[nE]: bit0 MPS>thr0 (CSDM-15), bit1 MPS>thr1 (CSDM-30), bit2 MPSR>thr2, bit3 MPSxSR>thr3, bit4 in the
 * MPS-95 element list, bit5 in the MPSxSR-95 list, bit7 element takes part; ps[nE], psxsr[nE] = PS_Old and
 * PSxSRArray (0 for excluded elements); volumes[5] = reference-configuration volume with bit0, bit0&bit1, bit2,
 * bit3 set and of all participating elements (ex5.cpp:1043-1066, Elements.cpp:30-38).  Pointers may be NULL. */
int ftb200_injury_get(ftb200_ctx *ctx, double *scalars12, int *extreme_elems4, unsigned char *flags, double *ps,
                      double *psxsr, double *volumes5);
/* Per-step 95th-percentile values (needs ftb200_record_history before ftb200_injury_begin) */
int ftb200_injury_history(ftb200_ctx *ctx, long long first, long long count, double *mps95, double *mpsxsr95);
/* CalculateMaximumPrincipalStrain (CalculateStrain.cpp:8-75) of every element for the displacements now on the
 * device: smax, smin, shear [nE] each (caller's element order), volume0[nE] = calculateVolume(e)
 * (CalculateCentroidAndVolume.cpp:26-37).  Pointers may be NULL. */
int ftb200_principal_strains(ftb200_ctx *ctx, double *smax, double *smin, double *shear, double *volume0);

/* ---- rigid-body prescribed-motion boundary condition of the brain drivers -------------
 * Replaces the drivers' ApplyAccBoundaryConditions callback (examples/ex5/ex5.cpp:339-371) and the node selection
 * of InitBoundaryCondition (:819-911) in the resident loop.  sizes[k], t[k], v[k], k = 0..5: the acceleration traces
 * angular x,y,z [rad/s^2] then linear x,y,z [m/s^2] over time [s] (ex5.cpp:92-98, already converted).  boundaryID
 * [boundarySize]: nodes that follow the rigid motion; NULL = every node of an element whose part has material 0
 * (ex5.cpp:819-846).  Those nodes get boundary = 1 on all dofs and u = v = a = 0; the 12 integrator states start at 0
 * (:897-900).  Each step the device advances the states with the Dormand-Prince step of odeint's runge_kutta_dopri5
 * (from Time - dt over dt) and sets u, v, a of the nodes from the rotation quaternion (math.cpp:122-158).
 * Call after ftb200_shape_functions and before ftb200_explicit_begin; combines with ftb200_set_bc for other nodes. */
int ftb200_set_rigid_bc(ftb200_ctx *ctx, const int sizes[6], const double *const t[6], const double *const v[6],
                        const int *boundaryID, int boundarySize);
/* yInt[12] = omega, r, v, d and ydotInt[12] (ex5.cpp:105); boundary_count = number of rigid-motion nodes */
int ftb200_get_rigid_state(ftb200_ctx *ctx, double *y12, double *ydot12, int *boundary_count);

#ifdef __cplusplus
}
#endif
#endif
