"""ctypes binding of the CPU oracle (oracle/femtech_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(femtech_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OracleState(C.Structure):
    """Mirror of `oracle_state` in femtech_oracle.h (same field order)."""
    _fields_ = [
        ("nNodes", C.c_int), ("nElements", C.c_int), ("nPID", C.c_int),
        ("coordinates", _dp), ("connectivity", _ip), ("pid", _ip),
        ("materialID", _ip), ("properties", _dp),
        ("shp", _dp), ("dshp", _dp), ("detJacobian", _dp), ("gaussWeights", _dp),
        ("F", _dp), ("detF", _dp), ("pk2", _dp),
        ("Hn_1", _dp), ("Hn_2", _dp), ("S0n", _dp),
        ("displacements", _dp), ("velocities", _dp), ("velocities_half", _dp),
        ("accelerations", _dp), ("mass", _dp), ("fe", _dp), ("fi", _dp), ("f_net", _dp),
        ("displacements_prev", _dp), ("accelerations_prev", _dp), ("fi_prev", _dp), ("fe_prev", _dp),
        ("boundary", _ip),
        ("world_rank", C.c_int), ("sendProcessCount", C.c_int),
        ("sendProcessID", _ip), ("sendNeighbourCountCum", _ip), ("sendNodeIndex", _ip),
        ("Time", C.c_double), ("dt", C.c_double), ("Wint_n", C.c_double), ("Wext_n", C.c_double),
        ("gpoff", _ip),
    ]


class OracleInjury(C.Structure):
    """Mirror of `oracle_injury` in femtech_oracle.h (same field order)."""
    _fields_ = [
        ("nElementsInjury", C.c_int), ("elementIDInjury", _ip),
        ("MPSgt15", _ip), ("MPSgt30", _ip), ("MPSRgt120", _ip), ("MPSxSRgt28", _ip),
        ("PS_Old", _dp), ("PSxSRArray", _dp),
        ("maxElemListMPS95", _ip), ("maxElemListMPSxSR95", _ip),
        ("maxElemCountMPS95", C.c_int), ("maxElemCountMPSxSR95", C.c_int),
        ("maxStrain", C.c_double), ("minStrain", C.c_double), ("maxShear", C.c_double), ("maxPSxSR", C.c_double),
        ("maxElem", C.c_int), ("minElem", C.c_int), ("shearElem", C.c_int), ("maxElemPSxSR", C.c_int),
        ("maxT", C.c_double), ("minT", C.c_double), ("maxShearT", C.c_double), ("maxTimePSxSR", C.c_double),
        ("maxMPS95", C.c_double), ("maxTimeMPS95", C.c_double), ("maxMPSxSR95", C.c_double),
        ("maxTimeMPSxSR95", C.c_double),
    ]


class OracleRigid(C.Structure):
    """Mirror of `oracle_rigid` in femtech_oracle.h."""
    _fields_ = [("size", C.c_int * 6), ("t", _dp * 6), ("v", _dp * 6), ("y", C.c_double * 12), ("ydot", C.c_double * 12),
                ("boundarySize", C.c_int), ("boundaryID", _ip)]


def build(fast=False):
    """Compile the oracle shared library if missing/stale; return its path."""
    name = "libfemtech_oracle_fast.so" if fast else "libfemtech_oracle.so"
    so = os.path.join(HERE, name)
    src = os.path.join(HERE, "femtech_oracle.c")
    hdr = os.path.join(HERE, "femtech_oracle.h")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", HERE, name], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        L = C.CDLL(build(fast))
        sp = C.POINTER(OracleState)
        L.oracle_ShapeFunctions.argtypes = [sp]
        L.oracle_AssembleLumpedMass_local.argtypes = [sp]
        L.oracle_GetForce_local.argtypes = [sp]
        L.oracle_GetForce_local.restype = C.c_int
        L.oracle_GetForce_finish.argtypes = [sp]
        L.oracle_CalculateAccelerations.argtypes = [sp]
        L.oracle_StableTimeStep_local.argtypes = [sp]
        L.oracle_StableTimeStep_local.restype = C.c_double
        L.oracle_CheckEnergy_local.argtypes = [sp, _dp]
        L.oracle_halo_sum.argtypes = [C.POINTER(sp), C.c_int, C.c_int]
        L.oracle_run_explicit.argtypes = [C.POINTER(sp), C.c_int, C.POINTER(_ip), _dp, C.c_double, C.c_int,
                                          C.c_double, C.c_double, C.c_int, _dp, _dp]
        L.oracle_run_explicit.restype = C.c_int
        L.oracle_volumeHexahedron.argtypes = [_dp]
        L.oracle_volumeHexahedron.restype = C.c_double
        L.oracle_areaHexahedronFace.argtypes = [_dp, _ip]
        L.oracle_areaHexahedronFace.restype = C.c_double
        L.oracle_CalculateTimeStep.argtypes = [sp, C.c_int]
        L.oracle_CalculateTimeStep.restype = C.c_double
        L.oracle_CalculateStrain.argtypes = [sp, _dp]
        qp = C.POINTER(OracleInjury)
        L.oracle_CalculateMaximumPrincipalStrain.argtypes = [sp, C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_InitInjuryCriterion.argtypes = [sp, qp, _ip, C.c_int]
        L.oracle_compute95thPercentileValue.argtypes = [C.POINTER(_dp), _ip, C.c_int]
        L.oracle_compute95thPercentileValue.restype = C.c_double
        L.oracle_CalculateInjuryCriterions.argtypes = [C.POINTER(sp), C.POINTER(qp), C.c_int, C.c_double, C.c_double]
        L.oracle_injury_volumes.argtypes = [sp, qp, _dp]
        L.oracle_run_explicit_injury.argtypes = [C.POINTER(sp), C.c_int, C.POINTER(_ip), _dp, C.c_double, C.c_int,
                                                 C.c_double, C.c_double, C.c_int, _dp, _dp, C.POINTER(qp)]
        L.oracle_run_explicit_injury.restype = C.c_int
        rp = C.POINTER(OracleRigid)
        L.oracle_interpolateLinear.argtypes = [C.c_int, _dp, _dp, C.c_double]
        L.oracle_interpolateLinear.restype = C.c_double
        L.oracle_computeDerivatives.argtypes = [rp, _dp, _dp, C.c_double]
        L.oracle_dopri5_step.argtypes = [rp, _dp, _dp, C.c_double, C.c_double]
        L.oracle_InitRigidBoundary.argtypes = [sp, rp]
        L.oracle_ApplyAccBoundaryConditions.argtypes = [sp, rp, C.c_double, C.c_double]
        L.oracle_run_explicit_rigid.argtypes = [sp, rp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp, qp]
        L.oracle_run_explicit_rigid.restype = C.c_int
        _libs[fast] = L
    return _libs[fast]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class OracleModel:
    """One rank's model: owns the numpy arrays the C struct points at."""

    NODAL = ["displacements", "velocities", "velocities_half", "accelerations", "mass", "fe", "fi", "f_net",
             "displacements_prev", "accelerations_prev", "fi_prev", "fe_prev"]

    def __init__(self, coordinates, connectivity, pid, materialID, properties, comm=None, world_rank=0, fast=False,
                 eptr=None):
        """eptr given: mixed C3D8 / C3D4 mesh, connectivity packed 8 or 4 node ids per element."""
        self.L = lib(fast)
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float64).reshape(-1)
        self.connectivity = np.ascontiguousarray(connectivity, dtype=np.int32).reshape(-1)
        self.pid = np.ascontiguousarray(pid, dtype=np.int32)
        self.materialID = np.ascontiguousarray(materialID, dtype=np.int32)
        self.properties = np.ascontiguousarray(properties, dtype=np.float64).reshape(-1)
        nN = self.coordinates.size // 3
        self.gpoff = None
        if eptr is None:
            nE = self.connectivity.size // 8
            nGP, nShp = 8 * nE, 64 * nE
        else:
            eptr = np.asarray(eptr, dtype=np.int64)
            nE = eptr.size - 1
            nen = np.diff(eptr)
            assert np.all((nen == 8) | (nen == 4)) and eptr[-1] == self.connectivity.size
            ngp = np.where(nen == 8, 8, 1)
            self.gpoff = np.concatenate([[0], np.cumsum(ngp)]).astype(np.int32)
            nGP, nShp = int(self.gpoff[-1]), int(np.sum(nen * ngp))
        self.nNodes, self.nElements = nN, nE
        self.shp = np.zeros(nShp)
        self.dshp = np.zeros(3 * nShp)
        self.detJacobian = np.zeros(nGP)
        self.gaussWeights = np.zeros(nGP)
        self.F = np.zeros(9 * nGP)
        self.detF = np.zeros(nGP)
        self.pk2 = np.zeros(6 * nGP)
        visco = bool(np.any(self.materialID == 5))
        self.Hn_1 = np.zeros(9 * nGP) if visco else None
        self.Hn_2 = np.zeros(9 * nGP) if visco else None
        self.S0n = np.zeros(9 * nGP) if visco else None
        for n in self.NODAL:
            setattr(self, n, np.zeros(3 * nN))
        self.boundary = np.zeros(3 * nN, dtype=np.int32)
        if comm is None:
            comm = dict(sendProcessID=np.zeros(0, np.int32), sendNeighbourCountCum=np.zeros(1, np.int32),
                        sendNodeIndex=np.zeros(0, np.int32))
        self.sendProcessID = np.ascontiguousarray(comm["sendProcessID"], dtype=np.int32)
        self.sendNeighbourCountCum = np.ascontiguousarray(comm["sendNeighbourCountCum"], dtype=np.int32)
        self.sendNodeIndex = np.ascontiguousarray(comm["sendNodeIndex"], dtype=np.int32)
        s = OracleState()
        s.nNodes, s.nElements, s.nPID = nN, nE, self.materialID.size
        s.coordinates, s.connectivity, s.pid = _d(self.coordinates), _i(self.connectivity), _i(self.pid)
        s.materialID, s.properties = _i(self.materialID), _d(self.properties)
        for n in ["shp", "dshp", "detJacobian", "gaussWeights", "F", "detF", "pk2"] + self.NODAL:
            setattr(s, n, _d(getattr(self, n)))
        if visco:
            s.Hn_1, s.Hn_2, s.S0n = _d(self.Hn_1), _d(self.Hn_2), _d(self.S0n)
        s.boundary = _i(self.boundary)
        s.world_rank = world_rank
        s.sendProcessCount = self.sendProcessID.size
        s.sendProcessID, s.sendNeighbourCountCum = _i(self.sendProcessID), _i(self.sendNeighbourCountCum)
        s.sendNodeIndex = _i(self.sendNodeIndex)
        s.Time = s.dt = s.Wint_n = s.Wext_n = 0.0
        s.gpoff = _i(self.gpoff) if self.gpoff is not None else None
        self.s = s

    # --- reference-named entry points -------------------------------------
    def ShapeFunctions(self):
        self.L.oracle_ShapeFunctions(C.byref(self.s))

    def AssembleLumpedMass(self):
        self.mass[:] = 0.0
        self.L.oracle_AssembleLumpedMass_local(C.byref(self.s))

    def GetForce(self):
        bad = self.L.oracle_GetForce_local(C.byref(self.s))
        self.L.oracle_GetForce_finish(C.byref(self.s))
        return bad

    def CalculateAccelerations(self):
        self.L.oracle_CalculateAccelerations(C.byref(self.s))

    def StableTimeStep(self):
        return self.L.oracle_StableTimeStep_local(C.byref(self.s))

    def CheckEnergy(self):
        out = np.zeros(3)
        self.L.oracle_CheckEnergy_local(C.byref(self.s), _d(out))
        return out

    def CalculateStrain(self):
        E = np.zeros(9 * self.nElements)
        self.L.oracle_CalculateStrain(C.byref(self.s), _d(E))
        return E

    @property
    def Time(self):
        return self.s.Time

    @property
    def dt(self):
        return self.s.dt


class InjuryCriteria:
    """ex5.cpp:62-83,1251-1306 state of one rank, arrays owned here."""

    def __init__(self, model, exclude_pids=()):
        n = model.nElements
        self.model = model
        self.q = OracleInjury()
        self.arrays = {}
        for name in ("elementIDInjury", "MPSgt15", "MPSgt30", "MPSRgt120", "MPSxSRgt28", "maxElemListMPS95",
                     "maxElemListMPSxSR95"):
            self.arrays[name] = np.zeros(max(n, 1), dtype=np.int32)
            setattr(self.q, name, _i(self.arrays[name]))
        for name in ("PS_Old", "PSxSRArray"):
            self.arrays[name] = np.zeros(max(n, 1))
            setattr(self.q, name, _d(self.arrays[name]))
        ex = np.ascontiguousarray(exclude_pids, dtype=np.int32)
        model.L.oracle_InitInjuryCriterion(C.byref(model.s), C.byref(self.q), _i(ex), len(ex))

    @property
    def n(self):
        return self.q.nElementsInjury

    def get(self, name):
        return self.arrays[name][:self.n]

    def principal(self, e):
        out = np.zeros(3)
        E = np.zeros(9)
        self.model.L.oracle_CalculateMaximumPrincipalStrain(C.byref(self.model.s), int(e), _d(E), _d(out[0:1]),
                                                            _d(out[1:2]), _d(out[2:3]))
        return out, E

    def volumes(self):
        out = np.zeros(5)
        self.model.L.oracle_injury_volumes(C.byref(self.model.s), C.byref(self.q), _d(out))
        return out

    def lists(self):
        return (self.arrays["maxElemListMPS95"][:self.q.maxElemCountMPS95].copy(),
                self.arrays["maxElemListMPSxSR95"][:self.q.maxElemCountMPSxSR95].copy())

    def scalars(self):
        q = self.q
        return np.array([q.maxStrain, q.maxT, q.minStrain, q.minT, q.maxShear, q.maxShearT, q.maxPSxSR, q.maxTimePSxSR,
                         q.maxMPS95, q.maxTimeMPS95, q.maxMPSxSR95, q.maxTimeMPSxSR95])

    def extreme_elems(self):
        q = self.q
        return np.array([q.maxElem, q.minElem, q.shearElem, q.maxElemPSxSR], dtype=np.int32)


def percentile95(arrays):
    """math.cpp:160-199 over the union of the per-rank arrays."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
    ptrs = (_dp * len(arrs))(*[_d(a) for a in arrs])
    sizes = np.array([len(a) for a in arrs], dtype=np.int32)
    return lib().oracle_compute95thPercentileValue(ptrs, _i(sizes), len(arrs))


def injury_step(models, injuries, Time, dt):
    P = len(models)
    sp, qp = C.POINTER(OracleState), C.POINTER(OracleInjury)
    arr = (sp * P)(*[C.pointer(m.s) for m in models])
    qarr = (qp * P)(*[C.pointer(i.q) for i in injuries])
    models[0].L.oracle_CalculateInjuryCriterions(arr, qarr, P, Time, dt)


class RigidBody:
    """ex5.cpp:92-106: acceleration traces (6 x (t, v), angular x,y,z then linear x,y,z) + the integrator state."""

    def __init__(self, model, tables):
        self.model = model
        self.rb = OracleRigid()
        self.keep = []
        for k, (t, v) in enumerate(tables):
            t = np.ascontiguousarray(t, dtype=np.float64)
            v = np.ascontiguousarray(v, dtype=np.float64)
            self.keep += [t, v]
            self.rb.size[k] = len(t)
            self.rb.t[k] = _d(t)
            self.rb.v[k] = _d(v)
        self.bid = np.zeros(max(model.nNodes, 1), dtype=np.int32)
        self.rb.boundaryID = _i(self.bid)
        model.L.oracle_InitRigidBoundary(C.byref(model.s), C.byref(self.rb))

    @property
    def boundaryID(self):
        return self.bid[:self.rb.boundarySize]

    @property
    def y(self):
        return np.array(self.rb.y[:])

    @property
    def ydot(self):
        return np.array(self.rb.ydot[:])

    def step(self, t, dt):
        """one dopri5 step of the 12 states from t to t + dt (ex5.cpp:344)"""
        y, yd = np.array(self.rb.y[:]), np.array(self.rb.ydot[:])
        self.model.L.oracle_dopri5_step(C.byref(self.rb), _d(y), _d(yd), t, dt)
        for i in range(12):
            self.rb.y[i], self.rb.ydot[i] = y[i], yd[i]


def run_explicit_rigid(model, rigid, tMax, maxSteps, reduction=0.8, failure_dt=1e-11, first_call=True, injury=None):
    """The ex5 loop (ex5.cpp:159-295), single rank.  Returns (steps, dt_hist, energy_hist)."""
    dth, eh = np.zeros(max(maxSteps, 1)), np.zeros(4 * max(maxSteps, 1))
    n = model.L.oracle_run_explicit_rigid(C.byref(model.s), C.byref(rigid.rb), tMax, maxSteps, reduction, failure_dt,
                                          1 if first_call else 0, _d(dth), _d(eh),
                                          C.byref(injury.q) if injury is not None else None)
    if n >= 0:
        return n, dth[:n].copy(), eh[:4 * n].reshape(n, 4).copy()
    return n, None, None


def halo_sum(models, field):
    """field: 'fi' or 'mass' (GetForce_3D.cpp:54-102 / Mass3D.cpp:77-125)."""
    sp = C.POINTER(OracleState)
    arr = (sp * len(models))(*[C.pointer(m.s) for m in models])
    models[0].L.oracle_halo_sum(arr, len(models), 0 if field == "fi" else 1)


def run_explicit(models, bc_kinds, bc_rate, tMax, maxSteps, reduction=0.8, failure_dt=1e-11, first_call=True,
                 record=True, injuries=None):
    """Benchmarking-Parallel.cpp:83-171 on P emulated ranks.  Returns
    (steps_or_negative_code, dt_hist, energy_hist[steps,4])."""
    P = len(models)
    sp = C.POINTER(OracleState)
    arr = (sp * P)(*[C.pointer(m.s) for m in models])
    kinds = [np.ascontiguousarray(k, dtype=np.int32) for k in bc_kinds]
    karr = (_ip * P)(*[_i(k) for k in kinds])
    rate = np.ascontiguousarray(bc_rate, dtype=np.float64)
    dth = np.zeros(max(maxSteps, 1)) if record else None
    eh = np.zeros(4 * max(maxSteps, 1)) if record else None
    qarr = None
    if injuries is not None:
        qp = C.POINTER(OracleInjury)
        qarr = (qp * P)(*[C.pointer(i.q) for i in injuries])
    n = models[0].L.oracle_run_explicit_injury(arr, P, karr, _d(rate), tMax, maxSteps, reduction, failure_dt,
                                               1 if first_call else 0, _d(dth) if record else None,
                                               _d(eh) if record else None, qarr)
    if record and n >= 0:
        return n, dth[:n].copy(), eh[:4 * n].reshape(n, 4).copy()
    return n, None, None


def read_ref_dump(path):
    """Parse a ref_dump record file (oracle/ref/ref_dump.cpp) into a dict."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        name = data[off:off + 32].split(b"\0")[0].decode()
        dtype = data[off + 32:off + 40].split(b"\0")[0].decode()
        count = int(np.frombuffer(data, dtype=np.int64, count=1, offset=off + 40)[0])
        off += 48
        dt = np.dtype("<" + dtype)
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=off).copy()
        off += count * dt.itemsize
    return out
