"""Host-side mirror of the reference's C-style API for the explicit step.

`FemTech` keeps the reference's global arrays as attributes with the same names,
layouts and ownership rules (include/GlobalVariables.h:18-127: AoS nodal arrays,
int boundary flags, hex8 connectivity of local node ids) and exposes the
reference's call sequence under the reference's names:

    ShapeFunctions()            src/fem/ShapeFunctions/ShapeFunctions.cpp:32-255
    AssembleLumpedMass()        src/fem/Mass/Mass3D.cpp:127-157
    StableTimeStep()            src/timestep/StableTimeStep.cpp:4-40
    GetForce()                  src/fem/SolidMechanics/GetForce.cpp:12-25 / GetForce_3D.cpp:5-53
    CalculateAccelerations()    src/fem/SolidMechanics/CalculateAcclerations.cpp:4-13
    CheckEnergy(time, flag)     src/fem/SolidMechanics/CheckEnergy.cpp:3-85
    CalculateStrain()           src/elements/ElementCalculations/CalculateStrain.cpp:77-97
    ExplicitDynamics(tFinal)    include/FemTech.h:50 (a stub in the reference; here: the whole
                                Benchmarking-Parallel.cpp:83-171 loop resident on the GPU)

Two modes, as in DESIGN.md: the *legacy* methods move the host arrays across
PCIe on every call (strict drop-in for the shipped drivers' host loops); the
*resident* ExplicitDynamics keeps state in HBM.  All arithmetic happens in the
CUDA library behind include/ftb200.h; this module only marshals pointers.
Fatal conditions raise FemTechB200Error carrying the reference's
TerminateFemTech code (1 unknown material, 3 bad input, 12 allocation, 19 time
step below FailureTimeStep).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FemTechB200Error

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class FemTech:
    """One rank's model (the reference: one MPI rank's globals)."""

    def __init__(self, coordinates, connectivity, pid, materialID, properties, comm=None, world_rank=0,
                 world_size=1, device=0, ExplicitTimeStepReduction=0.8, FailureTimeStep=1e-11, eptr=None):
        """eptr given: mixed C3D8 / C3D4 mesh, connectivity packed 8 or 4 node ids per element (GlobalVariables.h:32-33)."""
        self.L = _lib.load()
        self._h = C.c_void_p()
        self.world_rank, self.world_size = world_rank, world_size
        self.ExplicitTimeStepReduction = ExplicitTimeStepReduction
        self.FailureTimeStep = FailureTimeStep
        rc = self.L.ftb200_create(world_rank, world_size, device, C.byref(self._h))
        if rc:
            self._h = None
            raise FemTechB200Error(rc, "ftb200_create failed (no CUDA device? there is no CPU fallback)")
        # --- the reference's globals -----------------------------------------------------------
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float64).reshape(-1)
        self.connectivity = np.ascontiguousarray(connectivity, dtype=np.int32).reshape(-1)
        self.pid = np.ascontiguousarray(pid, dtype=np.int32)
        self.materialID = np.ascontiguousarray(materialID, dtype=np.int32)
        self.properties = np.ascontiguousarray(properties, dtype=np.float64).reshape(-1)
        self.nNodes = self.coordinates.size // 3
        self.nDOF = 3 * self.nNodes
        self.ndim = 3
        if eptr is None:
            self.nelements = self.connectivity.size // 8
            self.eptr = 8 * np.arange(self.nelements + 1, dtype=np.int32)
            if self.connectivity.size != 8 * self.nelements or self.pid.size != self.nelements:
                raise FemTechB200Error(3, "connectivity/pid size mismatch (pass eptr for meshes that are not all C3D8)")
            self._check(self.L.ftb200_upload_mesh(self._h, _d(self.coordinates), _i(self.connectivity), _i(self.pid),
                                                  self.nNodes, self.nelements))
        else:
            self.eptr = np.ascontiguousarray(eptr, dtype=np.int32)
            self.nelements = self.eptr.size - 1
            if self.pid.size != self.nelements or self.eptr[-1] != self.connectivity.size:
                raise FemTechB200Error(3, "connectivity/eptr/pid size mismatch")
            self._check(self.L.ftb200_upload_mesh_mixed(self._h, _d(self.coordinates), _i(self.connectivity), _i(self.eptr),
                                                        _i(self.pid), self.nNodes, self.nelements))
        self._check(self.L.ftb200_upload_materials(self._h, _i(self.materialID), _d(self.properties),
                                                   self.materialID.size))
        if comm is not None:
            self.sendProcessID = np.ascontiguousarray(comm["sendProcessID"], dtype=np.int32)
            self.sendNeighbourCountCum = np.ascontiguousarray(comm["sendNeighbourCountCum"], dtype=np.int32)
            self.sendNodeIndex = np.ascontiguousarray(comm["sendNodeIndex"], dtype=np.int32)
        else:
            self.sendProcessID = np.zeros(0, np.int32)
            self.sendNeighbourCountCum = np.zeros(1, np.int32)
            self.sendNodeIndex = np.zeros(0, np.int32)
        self.sendProcessCount = self.sendProcessID.size
        self._check(self.L.ftb200_upload_comm(self._h, self.sendProcessCount, _i(self.sendProcessID),
                                              _i(self.sendNeighbourCountCum), _i(self.sendNodeIndex)))
        # AllocateArrays() (src/fem/AllocateArrays.cpp:29-153): calloc'd nodal arrays
        for name in ("displacements", "velocities", "velocities_half", "accelerations", "fe", "fe_prev", "fi",
                     "fi_prev", "f_net", "displacements_prev", "accelerations_prev"):
            setattr(self, name, np.zeros(self.nDOF))
        self.boundary = np.zeros(self.nDOF, dtype=np.int32)
        self.mass = None
        self.Time = 0.0
        self.dt = 0.0
        self._Wint_n = 0.0  # CheckEnergy.cpp:4-5 function statics
        self._Wext_n = 0.0
        self.min_detJ = None

    # ---------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise FemTechB200Error(rc, self.L.ftb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.L.ftb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (int handle, e.g. torch.cuda.Stream().cuda_stream)."""
        self._check(self.L.ftb200_set_stream(self._h, C.c_void_p(cuda_stream)))

    @property
    def gpu_launches(self):
        return int(self.L.ftb200_launch_count(self._h))

    @property
    def affine_elements(self):
        """Hexahedra integrated by the parallelepiped kernel (after ShapeFunctions)."""
        return int(self.L.ftb200_affine_element_count(self._h))

    @property
    def brick_info(self):
        """Brick decomposition of the brick-fused step (after ShapeFunctions): dict, bricks = 0 when the two-kernel step is used."""
        out = (C.c_longlong * 8)()
        self._check(self.L.ftb200_brick_info(self._h, out))
        return {"bricks": int(out[0]), "interior_nodes": int(out[1]), "surface_nodes": int(out[2]), "partial_slots": int(out[3]),
                "dims": (int(out[4]), int(out[5]), int(out[6])), "active": bool(out[7])}

    def brick_maps(self):
        """(brick of every element, brick that finishes every node or -1 for a surface node), caller's numbering."""
        be = np.zeros(self.nelements, dtype=np.int32)
        bn = np.zeros(self.nNodes, dtype=np.int32)
        self._check(self.L.ftb200_brick_maps(self._h, be.ctypes.data_as(_ip), bn.ctypes.data_as(_ip)))
        return be, bn

    # --- one-time setup ----------------------------------------------------------------------------
    def ShapeFunctions(self):
        md = C.c_double()
        self._check(self.L.ftb200_shape_functions(self._h, C.byref(md)))
        self.min_detJ = md.value

    def AssembleLumpedMass(self):
        """Local lumped mass.  With world_size > 1 the caller sums the shared nodes afterwards
        (femtech_b200.dist.halo_sum_mass), as updateMassMatrixNeighbour does (Mass3D.cpp:77-125)."""
        self.mass = np.zeros(self.nDOF)
        self._check(self.L.ftb200_lumped_mass(self._h, _d(self.mass)))

    def refresh_mass(self):
        """Re-read `mass` from the device (after the neighbour sum of a multi-GPU run)."""
        self._check(self.L.ftb200_get_mass(self._h, _d(self.mass)))

    # --- legacy per-call path (host arrays in/out) ----------------------------------------------------
    def GetForce(self):
        fe = self.fe if np.any(self.fe) else None
        self._check(self.L.ftb200_get_force(self._h, _d(self.displacements), _d(fe), float(self.dt), _d(self.fi),
                                            _d(self.f_net)))

    def CalculateAccelerations(self):
        self._check(self.L.ftb200_calculate_accelerations(self._h, _i(self.boundary), _d(self.accelerations)))

    def StableTimeStep(self):
        out = C.c_double()
        self._check(self.L.ftb200_stable_time_step(self._h, _d(self.displacements), _i(self.boundary), C.byref(out)))
        dtMin = out.value
        if dtMin < self.FailureTimeStep:  # StableTimeStep.cpp:35-38
            raise FemTechB200Error(19, "Timestep too small, dt = %15.9e" % dtMin)
        return dtMin

    def CheckEnergy(self, time, writeFlag=1):
        """Returns (Wint_n, Wext_n, WKE, total) like the line CheckEnergy writes to energy_<uid>.dat."""
        out = np.zeros(3)
        fe = self.fe if np.any(self.fe) else None
        fep = self.fe_prev if np.any(self.fe_prev) else None
        self._check(self.L.ftb200_check_energy(self._h, _d(self.displacements), _d(self.displacements_prev),
                                               _d(self.velocities), _d(self.accelerations),
                                               _d(self.accelerations_prev), _d(self.fi), _d(self.fi_prev), _d(fe),
                                               _d(fep), _i(self.boundary), _d(out)))
        self._Wint_n += out[1]
        self._Wext_n += out[2]
        return self._Wint_n, self._Wext_n, out[0], abs(out[0] + self._Wint_n - self._Wext_n)

    def gp_outputs(self, F=True, detF=True, pk2=True, Eavg=False):
        """F[9 nGP], detF[nGP], pk2[6 nGP], Eavg[9E] of the last force evaluation, reference layouts (nGP = 8 per
        hexahedron + 1 per tetrahedron)."""
        nE = self.nelements
        nG = int(self.L.ftb200_gauss_point_count(self._h))
        res = {}
        if F:
            res["F"] = np.zeros(9 * nG)
        if detF:
            res["detF"] = np.zeros(nG)
        if pk2:
            res["pk2"] = np.zeros(6 * nG)
        if Eavg:
            res["Eavg"] = np.zeros(9 * nE)
        self._check(self.L.ftb200_get_gp_outputs(self._h, _d(res.get("F")), _d(res.get("detF")), _d(res.get("pk2")),
                                                 _d(res.get("Eavg"))))
        return res

    def CalculateStrain(self):
        return self.gp_outputs(F=False, detF=False, pk2=False, Eavg=True)["Eavg"]

    def CalculateMaximumPrincipalStrain(self, volume=False):
        """CalculateStrain.cpp:8-75 for every element at the displacements now on the device: (max, min, shear)
        arrays [nE] (+ calculateVolume(e) of the reference configuration when volume=True)."""
        nE = self.nelements
        a, b, c = np.zeros(nE), np.zeros(nE), np.zeros(nE)
        v = np.zeros(nE) if volume else None
        self._check(self.L.ftb200_principal_strains(self._h, _d(a), _d(b), _d(c), _d(v)))
        return (a, b, c, v) if volume else (a, b, c)

    # --- rigid-body prescribed motion of the brain drivers (ex5.cpp:339-371, :574-912) ----------------------------
    def set_rigid_bc(self, tables, boundaryID=None):
        """tables: six (t, v) traces -- angular acceleration x, y, z [rad/s^2] then linear x, y, z [m/s^2] over t [s].
        boundaryID None: every node of an element whose part has material 0 (InitBoundaryCondition)."""
        keep, sizes = [], np.zeros(6, dtype=np.int32)
        tp, vp = (_dp * 6)(), (_dp * 6)()
        for k, (t, v) in enumerate(tables):
            t = np.ascontiguousarray(t, dtype=np.float64)
            v = np.ascontiguousarray(v, dtype=np.float64)
            keep += [t, v]
            sizes[k] = len(t)
            tp[k], vp[k] = _d(t), _d(v)
        ids = None if boundaryID is None else np.ascontiguousarray(boundaryID, dtype=np.int32)
        self._check(self.L.ftb200_set_rigid_bc(self._h, _i(sizes), tp, vp, _i(ids), 0 if ids is None else len(ids)))

    def rigid_state(self):
        y, yd, n = np.zeros(12), np.zeros(12), C.c_int()
        self._check(self.L.ftb200_get_rigid_state(self._h, _d(y), _d(yd), C.byref(n)))
        return y, yd, n.value

    # --- injury criteria of the brain drivers (ex5.cpp:1251-1430), evaluated inside the resident loop ---------
    def InitInjuryCriterion(self, exclude_pids=(), thresholds=None):
        ex = np.ascontiguousarray(exclude_pids, dtype=np.int32)
        th = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float64)
        self._check(self.L.ftb200_injury_begin(self._h, _i(ex) if len(ex) else None, len(ex), _d(th)))

    def injury_end(self):
        self._check(self.L.ftb200_injury_end(self._h))

    def injury_results(self):
        """dict with the reference's names: scalars (maxStrain, maxT, minStrain, minT, maxShear, maxShearT,
        maxPSxSR, maxTimePSxSR, maxMPS95, maxTimeMPS95, maxMPSxSR95, maxTimeMPSxSR95), the four extreme elements,
        per-element flags / PS_Old / PSxSRArray, the MPS-95 element lists and the flagged volumes."""
        nE = self.nelements
        sc, el = np.zeros(12), np.zeros(4, dtype=np.int32)
        fl = np.zeros(nE, dtype=np.uint8)
        ps, px, vol = np.zeros(nE), np.zeros(nE), np.zeros(5)
        self._check(self.L.ftb200_injury_get(self._h, _d(sc), _i(el), fl.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(ps),
                                             _d(px), _d(vol)))
        incl = (fl & 0x80) != 0
        ids = np.nonzero(incl)[0].astype(np.int32)
        return dict(scalars=sc, extreme_elems=el, flags=fl, elementIDInjury=ids, PS_Old=ps[incl], PSxSRArray=px[incl],
                    MPSgt15=(fl[incl] & 1) != 0, MPSgt30=(fl[incl] & 2) != 0, MPSRgt120=(fl[incl] & 4) != 0,
                    MPSxSRgt28=(fl[incl] & 8) != 0, maxElemListMPS95=np.nonzero(fl & 16)[0].astype(np.int32),
                    maxElemListMPSxSR95=np.nonzero(fl & 32)[0].astype(np.int32), volumes=vol)

    def injury_history(self, first, count):
        a, b = np.zeros(count), np.zeros(count)
        self._check(self.L.ftb200_injury_history(self._h, int(first), int(count), _d(a), _d(b)))
        return a, b

    # --- resident path ---------------------------------------------------------------------------------
    def sync_in(self):
        self._check(self.L.ftb200_set_state(self._h, _d(self.displacements), _d(self.velocities),
                                            _d(self.accelerations), _i(self.boundary)))

    def sync_out(self, forces=True):
        self._check(self.L.ftb200_get_state(self._h, _d(self.displacements), _d(self.velocities),
                                            _d(self.accelerations), _i(self.boundary),
                                            _d(self.fi) if forces else None, _d(self.f_net) if forces else None))

    def set_bc(self, bc_kind, bc_rate):
        k = np.ascontiguousarray(bc_kind, dtype=np.int32)
        r = np.zeros(4)
        r[:len(bc_rate)] = bc_rate
        self._check(self.L.ftb200_set_bc(self._h, _i(k), _d(r)))

    def explicit_begin(self, energy_every=1, record_steps=0):
        """Step 0 of the drivers (Benchmarking-Parallel.cpp:83-91) on the device."""
        self.sync_in()
        if record_steps:
            self._check(self.L.ftb200_record_history(self._h, int(record_steps)))
        self._check(self.L.ftb200_explicit_begin(self._h, float(self.Time), float(self.ExplicitTimeStepReduction),
                                                 float(self.FailureTimeStep), int(energy_every)))
        self._poll()

    def _poll(self):
        steps, T, dt, st = C.c_longlong(), C.c_double(), C.c_double(), C.c_int()
        self._check(self.L.ftb200_explicit_poll(self._h, C.byref(steps), C.byref(T), C.byref(dt), C.byref(st)))
        self.Time, self.dt, self.steps_done, self.status_bits = T.value, dt.value, steps.value, st.value
        return steps.value

    def ExplicitDynamics(self, timeFinal, maxSteps=2 ** 62, sync=True):
        """The time loop on the device; returns the number of steps executed by this call."""
        steps, T, dt = C.c_longlong(), C.c_double(), C.c_double()
        rc = self.L.ftb200_explicit_run(self._h, float(timeFinal), int(maxSteps), C.byref(steps), C.byref(T),
                                        C.byref(dt))
        self.Time, self.dt = T.value, dt.value
        self._check(rc)
        if sync:
            self.sync_out()
        return steps.value

    def poll_async(self, out8_pinned):
        """Enqueue a D2H copy of (Time, dt, steps, status, Wint, Wext, WKE, balance) into pinned host memory; no sync."""
        self._check(self.L.ftb200_explicit_poll_async(self._h, _d(out8_pinned)))

    def step_ring(self, capacity):
        """Per-step records written by the device into pinned host memory (ftb200_step_ring): returns a live
        (capacity, 8) view -- Time, next dt, finished steps, status, Wint, Wext, WKE, |balance|; row (k - 1) % capacity
        belongs to step k once its column 2 reads k."""
        p = _dp()
        self._check(self.L.ftb200_step_ring(self._h, int(capacity), C.byref(p)))
        self._ring = np.ctypeslib.as_array(p, shape=(int(capacity), 8)) if capacity else None
        return self._ring

    def wait_step(self, k, timeout_s=120.0):
        """Spin on the step ring until the record of step k (counted from explicit_begin) has arrived; returns a copy."""
        import time as _t
        row = self._ring[(k - 1) % self._ring.shape[0]]
        t0 = _t.perf_counter()
        while row[2] != k:
            if row[2] > k:
                raise FemTechB200Error(3, "step ring overrun: record of step %d was overwritten" % k)
            if _t.perf_counter() - t0 > timeout_s:
                raise FemTechB200Error(100, "step %d did not arrive in the step ring" % k)
        return row.copy()

    def enable_partitioned_loop(self):
        """Single rank only: export this context's peer-memory window and import it as the only rank, so that run_async
        takes the loop of the multi-GPU runs (split element launches, dt through the window, k_adv_p2p) with zero
        neighbours.  bench.py uses it to separate the cost of that loop from the cost of scaling."""
        w = C.c_void_p()
        self._check(self.L.ftb200_p2p_export(self._h, None, C.byref(w)))
        arr = (C.c_void_p * 1)(w.value)
        self._check(self.L.ftb200_p2p_import(self._h, arr, 1, None, None, None))

    def run_async(self, timeFinal, steps):
        self._check(self.L.ftb200_explicit_run_async(self._h, float(timeFinal), int(steps)))

    def energy(self):
        out = np.zeros(4)
        self._check(self.L.ftb200_get_energy(self._h, _d(out)))
        return out

    def history(self, first, count):
        dth, eh = np.zeros(count), np.zeros(4 * count)
        self._check(self.L.ftb200_get_history(self._h, int(first), int(count), _d(dth), _d(eh)))
        return dth, eh.reshape(count, 4)

    def profile(self, on):
        self._check(self.L.ftb200_profile_enable(self._h, 1 if on else 0))

    def profile_get(self):
        e, n, ne, nn = C.c_double(), C.c_double(), C.c_longlong(), C.c_longlong()
        self._check(self.L.ftb200_profile_get(self._h, C.byref(e), C.byref(n), C.byref(ne), C.byref(nn)))
        return dict(elem_ms=e.value, node_ms=n.value, elem_launches=ne.value, node_launches=nn.value)


def measure_peaks(model, reps=5):
    """(fp64 TFLOP/s, copy GB/s) measured on the model's device (roofline denominators)."""
    a, b = C.c_double(), C.c_double()
    model._check(model.L.ftb200_measure_peaks(model._h, int(reps), C.byref(a), C.byref(b)))
    return a.value, b.value


def legacy_explicit_loop(model, bc_kind, bc_rate, tMax, maxSteps, record=False):
    """The shipped drivers' time loop (Benchmarking-Parallel.cpp:83-171) with the HOST loops of the
    driver done in numpy and the four library calls going through the legacy path -- the strict
    drop-in mode.  Used by the parity tests and bench.py's e2e leg."""
    m = model
    kind = np.asarray(bc_kind)
    bc = kind > 0
    rate = np.asarray(bc_rate, dtype=np.float64)[kind[bc]]

    def apply_bc():
        m.boundary[bc] = 1
        m.displacements[bc] = m.Time * rate
        m.velocities[bc] = rate
        m.accelerations[bc] = 0.0

    apply_bc()
    m.dt = m.ExplicitTimeStepReduction * m.StableTimeStep()
    m.GetForce()
    m.CalculateAccelerations()
    steps = 0
    dth, eh = [], []
    while m.Time < tMax and steps < maxSteps:
        t_n = m.Time
        t_np1 = m.Time + m.dt
        m.Time = t_np1
        t_nphalf = 0.5 * (t_np1 + t_n)
        free = m.boundary == 0
        dth.append(m.dt)
        m.velocities_half[:] = np.where(free, m.velocities + (t_nphalf - t_n) * m.accelerations, m.velocities)
        m.displacements_prev[:] = m.displacements
        m.accelerations_prev[:] = m.accelerations
        m.fi_prev[:] = m.fi
        m.fe_prev[:] = m.fe
        m.displacements[free] = m.displacements[free] + m.dt * m.velocities_half[free]
        apply_bc()
        m.GetForce()
        m.CalculateAccelerations()
        free = m.boundary == 0
        m.velocities[free] = m.velocities_half[free] + (t_np1 - t_nphalf) * m.accelerations[free]
        e = m.CheckEnergy(m.Time, 0)
        if record:
            eh.append(e)
        steps += 1
        m.dt = m.ExplicitTimeStepReduction * m.StableTimeStep()
    return steps, np.array(dth), np.array(eh)
