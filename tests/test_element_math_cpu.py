"""The per-element arithmetic the CUDA kernels run (femtech_b200/csrc/
hex8_element.cuh, __host__ __device__) compiled with g++ and compared with the
oracle on the golden end states: element forces, F, det F, PK2, stable dt and
lumped mass.  Runs without a GPU; it checks the algebra of the mode-basis
reformulation (which is not a transcription of the reference), not the kernels.
Tolerance: 1e-11 relative to the field maximum (pure rounding differences)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, rank_dict
from oracle import pyoracle as po

HARN = os.path.join(ROOT, "tests", "cpu_harness")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def harness():
    so = os.path.join(HARN, "libelem_harness.so")
    src = os.path.join(HARN, "elem_harness.cpp")
    hdr = os.path.join(ROOT, "femtech_b200", "csrc", "hex8_element.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", so, src])
    L = C.CDLL(so)
    L.harness_element.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp]
    L.harness_element.restype = C.c_int
    L.harness_element_affine.argtypes = L.harness_element.argtypes
    L.harness_element_affine.restype = C.c_int
    L.harness_element_affine_staged.argtypes = L.harness_element.argtypes
    L.harness_element_affine_staged.restype = C.c_int
    L.harness_element_affine_cj.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp]
    L.harness_element_affine_cj.restype = C.c_int
    L.harness_element_brick.argtypes = L.harness_element.argtypes
    L.harness_element_brick.restype = C.c_int
    L.harness_face_amax.argtypes = [_dp, _dp]
    L.harness_mass.argtypes = [_dp, C.c_double, _dp]
    L.harness_mass.restype = C.c_double
    L.harness_principal.argtypes = [_dp, _dp, _dp]
    L.harness_tet.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    L.harness_tet.restype = C.c_int
    return L


def part_params(materialID, properties, dt):
    """Host-side derived per-part block (mirrors femtech_b200/csrc: MP_* layout)."""
    props = np.asarray(properties, float).reshape(-1, 9)
    mp = np.zeros((props.shape[0], 16))
    mp[:, :9] = props
    for p in range(props.shape[0]):
        rho, mu, lam = props[p, 0], props[p, 1], props[p, 2]
        with np.errstate(all="ignore"):
            nu = 0.5 * lam / (lam + mu)
            mp[p, 9] = np.sqrt(lam * (1.0 / nu - 1.0) / rho)
        mp[p, 10] = lam + 2.0 * mu / 3.0
        if materialID[p] == 5 and dt > 0:
            rt1, rt2 = dt / props[p, 6], dt / props[p, 8]
            c11, c12 = np.exp(-rt1), np.exp(-rt2)
            mp[p, 11:15] = [c11, c12, props[p, 5] * (1 - c11) / rt1, props[p, 7] * (1 - c12) / rt2]
        mp[p, 15] = materialID[p]
    return mp


VOIGT = [0, 4, 8, 7, 6, 3]


@pytest.mark.parametrize("name", ["cube4j_m1", "cube4j_m2", "cube4j_m3", "cube4j_m4", "cube4j_m5", "cube6mix_p1"])
def test_element_force_F_pk2_dt_match_oracle(harness, name):
    g = golden(name)
    d = rank_dict(g, 0)
    m = po.OracleModel(d["coordinates"], d["connectivity"], d["pid"], d["materialID"], d["properties"])
    m.ShapeFunctions()
    # put the oracle into the golden end state, but with the history of the step BEFORE the
    # last GetForce: re-run the whole thing to steps-1, then do the last step by hand
    from femtech_b200 import mesh
    kind, rate = mesh.benchmark_bc(d["coordinates"], dMax=float(g["param_dMax"]), tMax=float(g["param_tMax"]))
    m.AssembleLumpedMass()
    nsteps = int(d["steps"][0])
    n, _, _ = po.run_explicit([m], [kind], rate, float(g["param_tMax"]), nsteps - 1)
    assert n == nsteps - 1
    # advance the displacement like the driver does, then evaluate forces both ways
    Time, dt = m.Time, m.dt
    t_np1 = Time + dt
    t_half = 0.5 * (t_np1 + Time)
    free = m.boundary == 0
    vh = np.where(free, m.velocities + (t_half - Time) * m.accelerations, m.velocities)
    m.displacements[free] += dt * vh[free]
    bc = kind > 0
    m.displacements[bc] = t_np1 * rate[kind[bc]]
    hist0 = None
    if m.Hn_1 is not None:
        hist0 = [m.Hn_1.copy(), m.Hn_2.copy(), m.S0n.copy()]
    m.GetForce()
    mp = part_params(d["materialID"], d["properties"], dt)
    X = m.coordinates.reshape(-1, 3)
    U = m.displacements.reshape(-1, 3)
    conn = m.connectivity.reshape(-1, 8)
    nE = conn.shape[0]
    fi = np.zeros_like(U)
    F = np.zeros(72 * nE); detF = np.zeros(8 * nE); pk2 = np.zeros(48 * nE); dts = np.zeros(nE)
    for e in range(nE):
        Xe = np.ascontiguousarray(X[conn[e]]).reshape(-1)
        Ue = np.ascontiguousarray(U[conn[e]]).reshape(-1)
        p = int(d["pid"][e])
        hist = np.zeros(144)
        if hist0 is not None:
            for gp in range(8):
                for a, arr in enumerate(hist0):
                    hist[18 * gp + 6 * a:18 * gp + 6 * a + 6] = arr[72 * e + 9 * gp + np.array(VOIGT)]
        fe = np.zeros(24); dte = np.zeros(1)
        mpp = np.ascontiguousarray(mp[p])
        st = harness.harness_element(Xe.ctypes.data_as(_dp), Ue.ctypes.data_as(_dp), int(d["materialID"][p]),
                                     mpp.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), 1, fe.ctypes.data_as(_dp),
                                     dte.ctypes.data_as(_dp), F[72 * e:].ctypes.data_as(_dp),
                                     detF[8 * e:].ctypes.data_as(_dp), pk2[48 * e:].ctypes.data_as(_dp))
        assert (st & ~4) == 0  # bit 4 (det F <= 0) is informational: StVK golden run inverts elements in the reference too
        np.add.at(fi, conn[e], fe.reshape(8, 3))
        dts[e] = dte[0]
        if hist0 is not None:  # updated history must match the oracle's
            for gp in range(8):
                for a, arr in enumerate([m.Hn_1, m.Hn_2, m.S0n]):
                    want = arr[72 * e + 9 * gp + np.array(VOIGT)]
                    got = hist[18 * gp + 6 * a:18 * gp + 6 * a + 6]
                    assert np.allclose(got, want, rtol=0, atol=1e-10 * max(np.abs(m.S0n).max(), 1e-300))

    def rel(a, b):
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

    assert rel(fi.reshape(-1), m.fi) < 1e-11
    assert rel(F, m.F) < 1e-13
    assert rel(detF, m.detF) < 1e-13
    assert rel(pk2, m.pk2) < 1e-10
    ref_dt = np.array([po.lib().oracle_CalculateTimeStep(C.byref(m.s), e) for e in range(nE)])
    assert np.allclose(dts, ref_dt, rtol=1e-12)


def test_lumped_mass_matches_oracle(harness):
    d = rank_dict(golden("cube4j_m1"), 0)
    X = d["coordinates"].reshape(-1, 3)
    conn = d["connectivity"].reshape(-1, 8)
    mass = np.zeros(X.shape[0])
    for e in range(conn.shape[0]):
        me = np.zeros(8)
        Xe = np.ascontiguousarray(X[conn[e]]).reshape(-1)
        det = harness.harness_mass(Xe.ctypes.data_as(_dp), float(d["properties"][0]), me.ctypes.data_as(_dp))
        assert det > 0
        np.add.at(mass, conn[e], me)
    assert np.allclose(np.repeat(mass, 3), d["mass"], rtol=1e-13)


@pytest.mark.parametrize("name", ["inj6_p1", "inj6b_p1"])
def test_principal_strains_match_reference(harness, name):
    """principal_strains + StrainSink (what k_elem<..., WITH_INJ> runs) vs the reference's own
    CalculateMaximumPrincipalStrain on the final state of the injury fixtures (PS_Old = max principal strain)."""
    g = golden(name)
    d = rank_dict(g, 0)
    X = d["coordinates"].reshape(-1, 3)
    U = d["displacements"].reshape(-1, 3)
    conn = d["connectivity"].reshape(-1, 8)
    Eavg = d["Eavg"].reshape(-1, 9)
    ps = dict(zip(d["inj_elems"].tolist(), d["inj_ps_old"].tolist()))
    worst = 0.0
    for e in range(conn.shape[0]):
        Xe = np.ascontiguousarray(X[conn[e]])
        Ue = np.ascontiguousarray(U[conn[e]])
        out = np.zeros(9)
        harness.harness_principal(Xe.ctypes.data_as(_dp), Ue.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        # E = (0.5/8) sum F^T F - 0.5 I against the reference's Eavg (column-major 3x3)
        Em = Eavg[e].reshape(3, 3)
        mine = 0.0625 * np.array([out[3], out[4], out[5], out[6], out[7], out[8]]) - np.array([0.5, 0.5, 0.5, 0, 0, 0])
        want = np.array([Em[0, 0], Em[1, 1], Em[2, 2], Em[1, 2], Em[0, 2], Em[0, 1]])
        assert np.max(np.abs(mine - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
        ev = np.linalg.eigvalsh(Em)
        assert out[0] == pytest.approx(max(ev[-1], 0.0), rel=1e-9, abs=1e-12)
        assert out[1] == pytest.approx(min(ev[0], 0.0), rel=1e-9, abs=1e-12)
        assert out[2] == pytest.approx(0.5 * (ev[-1] - ev[0]), rel=1e-9, abs=1e-12)
        if e in ps:
            worst = max(worst, abs(out[0] - ps[e]) / max(abs(ps[e]), 1e-12))
    assert worst <= 1e-9  # tolerance of the north star for strains


def test_principal_strains_double_root_is_finite(harness):
    """Uniaxial stretch along a skew axis: a double eigenvalue with non-zero off-diagonals.  The reference's acos
    argument may round past 1 there (NaN -> it reports 0); the device formula clamps and returns the eigenvalues."""
    n = np.array([1.0, 2.0, 2.0]) / 3.0
    lam = 1.3
    Fm = np.eye(3) + (lam - 1.0) * np.outer(n, n)
    g = golden("inj6_p1")
    conn = rank_dict(g, 0)["connectivity"].reshape(-1, 8)[0]
    X = rank_dict(g, 0)["coordinates"].reshape(-1, 3)[conn]
    X = np.ascontiguousarray(X)
    U = np.ascontiguousarray(X @ Fm.T - X)
    out = np.zeros(9)
    harness.harness_principal(X.ctypes.data_as(_dp), U.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    assert np.all(np.isfinite(out))
    assert out[0] == pytest.approx(0.5 * (lam * lam - 1.0), rel=1e-7)
    assert out[1] == 0.0 or abs(out[1]) < 1e-9


@pytest.mark.parametrize("name", ["mix4_p1", "mix4v_p1"])
def test_tet4_element_matches_oracle(harness, name):
    """tet4_element / tet4_lumped_mass (the C3D4 branch of the device kernels) vs the oracle on the tetrahedra of the mixed
    fixtures: one force evaluation from the state before the last step, incl. the Prony history of one-point elements."""
    from femtech_b200 import mesh
    g = golden(name)
    d = rank_dict(g, 0)
    m = po.OracleModel(d["coordinates"], d["connectivity"], d["pid"], d["materialID"], d["properties"], eptr=d["eptr"])
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    kind, rate = mesh.benchmark_bc(d["coordinates"], dMax=float(g["param_dMax"]), tMax=float(g["param_tMax"]))
    nsteps = int(d["steps"][0])
    n, _, _ = po.run_explicit([m], [kind], rate, float(g["param_tMax"]), nsteps - 1)
    assert n == nsteps - 1
    Time, dt = m.Time, m.dt
    t_np1 = Time + dt
    free = m.boundary == 0
    vh = np.where(free, m.velocities + (0.5 * (t_np1 + Time) - Time) * m.accelerations, m.velocities)
    m.displacements[free] += dt * vh[free]
    bc = kind > 0
    m.displacements[bc] = t_np1 * rate[kind[bc]]
    hist0 = [a.copy() for a in (m.Hn_1, m.Hn_2, m.S0n)] if m.Hn_1 is not None else None
    m.GetForce()
    mp = part_params(d["materialID"], d["properties"], dt)
    X, U = m.coordinates.reshape(-1, 3), m.displacements.reshape(-1, 3)
    eptr, gpoff = d["eptr"], m.gpoff
    ntet = 0
    masses = np.zeros(X.shape[0])
    for e in range(m.nElements):
        if eptr[e + 1] - eptr[e] != 4:
            continue
        ntet += 1
        nodes = m.connectivity[eptr[e]:eptr[e + 1]]
        Xe, Ue = np.ascontiguousarray(X[nodes]).reshape(-1), np.ascontiguousarray(U[nodes]).reshape(-1)
        p = int(d["pid"][e])
        g0 = int(gpoff[e])
        hist = np.zeros(18)
        if hist0 is not None:
            for a, arr in enumerate(hist0):
                hist[6 * a:6 * a + 6] = arr[9 * g0 + np.array(VOIGT)]
        fe, dte, F9, dF, pk, me = np.zeros(12), np.zeros(1), np.zeros(9), np.zeros(1), np.zeros(6), np.zeros(4)
        mpp = np.ascontiguousarray(mp[p])
        st = harness.harness_tet(Xe.ctypes.data_as(_dp), Ue.ctypes.data_as(_dp), int(d["materialID"][p]), mpp.ctypes.data_as(_dp),
                                 hist.ctypes.data_as(_dp), 1, fe.ctypes.data_as(_dp), dte.ctypes.data_as(_dp),
                                 F9.ctypes.data_as(_dp), dF.ctypes.data_as(_dp), pk.ctypes.data_as(_dp), me.ctypes.data_as(_dp))
        assert st == 0
        assert np.allclose(F9, m.F[9 * g0:9 * g0 + 9], rtol=0, atol=1e-13)
        assert abs(dF[0] - m.detF[g0]) <= 1e-13 * abs(m.detF[g0])
        scale = max(np.abs(m.pk2).max(), 1e-300)
        assert np.allclose(pk, m.pk2[6 * g0:6 * g0 + 6], rtol=0, atol=1e-10 * scale)
        assert dte[0] == pytest.approx(po.lib().oracle_CalculateTimeStep(C.byref(m.s), e), rel=1e-12)
        if hist0 is not None:
            for a, arr in enumerate([m.Hn_1, m.Hn_2, m.S0n]):
                assert np.allclose(hist[6 * a:6 * a + 6], arr[9 * g0 + np.array(VOIGT)], rtol=0, atol=1e-10 * max(np.abs(m.S0n).max(), 1e-300))
        np.add.at(masses, nodes, me)
    assert ntet > 50
    # forces: assemble tets through the harness and hexahedra through the hex harness == the oracle's fi
    fi = np.zeros_like(U)
    for e in range(m.nElements):
        nodes = m.connectivity[eptr[e]:eptr[e + 1]]
        Xe, Ue = np.ascontiguousarray(X[nodes]).reshape(-1), np.ascontiguousarray(U[nodes]).reshape(-1)
        p = int(d["pid"][e])
        g0 = int(gpoff[e])
        mpp = np.ascontiguousarray(mp[p])
        dte = np.zeros(1)
        if len(nodes) == 4:
            hist = np.zeros(18)
            if hist0 is not None:
                for a, arr in enumerate(hist0):
                    hist[6 * a:6 * a + 6] = arr[9 * g0 + np.array(VOIGT)]
            fe = np.zeros(12)
            harness.harness_tet(Xe.ctypes.data_as(_dp), Ue.ctypes.data_as(_dp), int(d["materialID"][p]), mpp.ctypes.data_as(_dp),
                                hist.ctypes.data_as(_dp), 1, fe.ctypes.data_as(_dp), dte.ctypes.data_as(_dp), np.zeros(9).ctypes.data_as(_dp),
                                np.zeros(1).ctypes.data_as(_dp), np.zeros(6).ctypes.data_as(_dp), None)
            np.add.at(fi, nodes, fe.reshape(4, 3))
        else:
            hist = np.zeros(144)
            if hist0 is not None:
                for gp in range(8):
                    for a, arr in enumerate(hist0):
                        hist[18 * gp + 6 * a:18 * gp + 6 * a + 6] = arr[9 * (g0 + gp) + np.array(VOIGT)]
            fe = np.zeros(24)
            harness.harness_element(Xe.ctypes.data_as(_dp), Ue.ctypes.data_as(_dp), int(d["materialID"][p]), mpp.ctypes.data_as(_dp),
                                    hist.ctypes.data_as(_dp), 1, fe.ctypes.data_as(_dp), dte.ctypes.data_as(_dp), np.zeros(72).ctypes.data_as(_dp),
                                    np.zeros(8).ctypes.data_as(_dp), np.zeros(48).ctypes.data_as(_dp))
            np.add.at(fi, nodes, fe.reshape(8, 3))
    assert np.abs(fi.reshape(-1) - m.fi).max() < 1e-11 * np.abs(m.fi).max()


def _call_elem(fn, X, U, mat, mp, hist):
    fe, dte, F, dF, pk = np.zeros(24), np.zeros(1), np.zeros(72), np.zeros(8), np.zeros(48)
    st = fn(np.ascontiguousarray(X).reshape(-1).ctypes.data_as(_dp), np.ascontiguousarray(U).reshape(-1).ctypes.data_as(_dp), mat,
            mp.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), 1, fe.ctypes.data_as(_dp), dte.ctypes.data_as(_dp),
            F.ctypes.data_as(_dp), dF.ctypes.data_as(_dp), pk.ctypes.data_as(_dp))
    return st, fe, dte[0], F, dF, pk


@pytest.mark.parametrize("mat", [1, 2, 3, 4, 5])
def test_affine_hexahedron_path_equals_general_path(harness, mat):
    """hex8_element_affine_in (parallelepiped elements: cof(J0), J0^-1 once per element) against the general mode-basis
    path on sheared parallelepipeds with dyadic coordinates (edge vectors bit-equal) and random displacements: forces,
    F, det F, PK2, dt and the updated Prony history agree to rounding.  A jittered element must be refused."""
    rng = np.random.default_rng(7 + mat)
    props = np.array([1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747])
    if mat in (1, 2, 3):
        props = np.array([1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0.0])
    mp = np.ascontiguousarray(part_params([mat], props, 1e-6)[0])
    signs = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]])
    for trial in range(20):
        # edge vectors and origin on a 2^-10 grid: every node coordinate and every edge difference is exact
        D = (np.eye(3) * 64 + rng.integers(-16, 17, size=(3, 3))) / 1024.0
        x0 = rng.integers(-2048, 2048, size=3) / 1024.0
        X = x0 + ((signs + 1) // 2) @ D
        U = 0.004 * rng.standard_normal((8, 3))
        h0 = 50.0 * rng.standard_normal(144)
        ha, hg = h0.copy(), h0.copy()
        sa, fa, dta, Fa, dFa, pka = _call_elem(harness.harness_element_affine, X, U, mat, mp, ha)
        sg, fg, dtg, Fg, dFg, pkg = _call_elem(harness.harness_element, X, U, mat, mp, hg)
        assert sa == sg == 0
        # the kernel's staging-slot plan replayed on the host gives the very same bits
        hs = h0.copy()
        ss, fs, dts_, Fs, dFs, pks = _call_elem(harness.harness_element_affine_staged, X, U, mat, mp, hs)
        assert ss == 0 and np.array_equal(fs, fa) and dts_ == dta and np.array_equal(Fs, Fa) and np.array_equal(hs, ha)
        # the brick kernel's variant (45 scratch slots, det J0 folded into the weights): rounding differences only
        hb = h0.copy()
        sb, fb, dtb, Fb, dFb, pkb = _call_elem(harness.harness_element_brick, X, U, mat, mp, hb)
        assert sb == 0 and np.array_equal(Fb, Fa) and np.array_equal(pkb, pka) and dtb == dta and np.array_equal(hb, ha)
        assert np.abs(fb - fa).max() <= 1e-14 * np.abs(fa).max()
        assert np.abs(Fa - Fg).max() < 1e-14
        assert np.abs(dFa - dFg).max() < 1e-14
        assert np.abs(fa - fg).max() <= 1e-13 * np.abs(fg).max()
        assert np.abs(pka - pkg).max() <= 1e-12 * max(np.abs(pkg).max(), 1.0)
        assert dta == pytest.approx(dtg, rel=1e-14)
        if mat == 5:
            assert np.abs(ha - hg).max() <= 1e-12 * max(np.abs(hg).max(), np.abs(pkg).max())  # history carries stress differences
            assert np.abs(ha - h0).max() > 0
    Xj = X + 1e-9 * rng.standard_normal((8, 3))
    assert _call_elem(harness.harness_element_affine, Xj, U, mat, mp, h0.copy())[0] == -1


@pytest.mark.parametrize("mat", [1, 4])
@pytest.mark.parametrize("strain", [0.0, 1e-9, 1e-5, 0.004, 0.05, 0.3])
def test_current_jacobian_form_equals_displacement_gradient_form(harness, strain, mat):
    """hex8_element_affine_cj (neo-Hookean: Q = mu Ft M + c cof Ft with the linear part summed over the Gauss points in
    closed form; HGO: Q = alpha Ft M + gamma cof Ft with tr B from the same product)
    against hex8_element_affine_in and the general path on sheared parallelepipeds, from the rest state (where the new form
    cancels to rounding instead of returning exact zeros) to 30 % displacement gradients: forces agree to 1e-14 of
    mu * (face area) -- the scale of the two terms that cancel -- plus 1e-13 of the force maximum; dt agrees to 1e-14,
    the status is the same and the staged input gives the same bits."""
    rng = np.random.default_rng(23)
    props = np.array([1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0.0])
    if mat == 4:
        props = np.array([1000.0, 2673.23, 2.189982178466e8, 25459.0, 10.0, 0, 0, 0, 0])
    mp = np.ascontiguousarray(part_params([mat], props, 1e-6)[0])
    signs = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]])
    for trial in range(20):
        D = (np.eye(3) * 64 + rng.integers(-16, 17, size=(3, 3))) / 1024.0
        x0 = rng.integers(-2048, 2048, size=3) / 1024.0
        X = x0 + ((signs + 1) // 2) @ D
        U = strain * 0.0625 * rng.standard_normal((8, 3))
        h0 = np.zeros(144)
        sa, fa, dta, Fa, dFa, pka = _call_elem(harness.harness_element_affine, X, U, mat, mp, h0)
        for staged in (0, 1):
            fn, dtn = np.zeros(24), np.zeros(1)
            st = harness.harness_element_affine_cj(np.ascontiguousarray(X).reshape(-1).ctypes.data_as(_dp),
                                                   np.ascontiguousarray(U).reshape(-1).ctypes.data_as(_dp), mat, mp.ctypes.data_as(_dp), staged,
                                                   fn.ctypes.data_as(_dp), dtn.ctypes.data_as(_dp))
            assert st == sa and (st & ~4) == 0  # bit 4: an inverted Gauss point (possible at the 30 % level), both forms report it
            if st:
                continue  # ln J of a negative J: NaN in both
            if staged:
                assert np.array_equal(fn, f_direct) and dtn[0] == dt_direct
            f_direct, dt_direct = fn.copy(), dtn[0]
            scale = (props[1] if mat == 1 else props[2]) * 0.0625 ** 2  # modulus * face area: the size of the terms that cancel
            # HGO at 30 %: exp(k2 Ea^2) ~ 1e80 amplifies the rounding of tr B by k2 Ea^2 ~ 200
            assert np.abs(fn - fa).max() <= 1e-14 * scale + (1e-13 if mat == 1 else 1e-11) * np.abs(fa).max()
            assert dtn[0] == pytest.approx(dta, rel=1e-14)
            assert abs(fn.reshape(8, 3).sum(axis=0)).max() <= 1e-15 * scale + 1e-15 * np.abs(fn).max()  # momentum balance of the mode basis


def test_filtered_face_maximum_equals_every_face_maximum(harness):
    """hex_face_amax (bounds first, square roots only for the faces that can be the largest) against hex_face_amax_all
    (every face integrated, the form pinned to the oracle above) on random hexahedra: regular, mildly and strongly
    distorted, tiny (all faces below the reference's 1e-6 threshold) and with near-tied faces."""
    rng = np.random.default_rng(11)
    base = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    out = np.zeros(2)
    worst = 0.0
    for trial in range(4000):
        h = 10.0 ** rng.uniform(-6.5, -1.0)                     # element size: around and far above the 1e-6 test
        aspect = np.where(rng.random(3) < 0.3, 1.0, rng.uniform(0.5, 2.0, 3))  # cubes (near-tied faces) and bricks
        jit = rng.choice([0.0, 1e-9, 1e-3, 0.05, 0.25])
        X = (base * aspect + jit * rng.uniform(-1, 1, (8, 3))) * h
        X = X @ np.linalg.qr(rng.standard_normal((3, 3)))[0].T if rng.random() < 0.5 else X
        X = X + rng.uniform(-1, 1, 3) * h
        Xc = np.ascontiguousarray(X).reshape(-1)
        harness.harness_face_amax(Xc.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        assert out[1] > 0
        worst = max(worst, abs(out[0] - out[1]) / out[1])
    assert worst <= 1e-13
