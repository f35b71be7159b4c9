"""Pins the CPU oracle (oracle/femtech_oracle.c) against the reference itself.

The fixtures in tests/golden/ were dumped from the UNMODIFIED reference sources
(oracle/make_golden.py -> oracle/_ref/ref_dump_exact).  The oracle restates the
reference operation by operation, so with -ffp-contract=off on both sides the
comparison is BIT-EXACT, not a tolerance.  Also checks the reference's own
known-answer test (examples/ex9/compareResults.py: 5 % vs Abaqus).
"""
import numpy as np
import pytest

from conftest import golden, rank_dict
from femtech_b200 import mesh
from oracle import pyoracle as po

SINGLE = ["ex9_1elt", "bench10_p1", "cube4j_m1", "cube4j_m2", "cube4j_m3", "cube4j_m4", "cube4j_m5", "cube6mix_p1",
          "cube4j_m1_1k", "cube4j_m4_1k", "cube4j_m5_1k"]  # _1k: 1000 steps (the north-star bar), oracle/make_golden.py group H
MULTI = ["bench10_p2", "bench10_p4", "bench10_p8", "cube6mix_p3"]


def _models(g):
    P = int(g["nranks"])
    models, kinds, rate = [], [], None
    for r in range(P):
        d = rank_dict(g, r)
        comm = None
        if "sendProcessID" in d:
            comm = {k: d[k] for k in ("sendProcessID", "sendNeighbourCountCum", "sendNodeIndex")}
        m = po.OracleModel(d["coordinates"], d["connectivity"], d["pid"], d["materialID"], d["properties"], comm=comm,
                           world_rank=r)
        k, rate = mesh.benchmark_bc(d["coordinates"], dMax=float(g["param_dMax"]), tMax=float(g["param_tMax"]))
        models.append(m)
        kinds.append(k)
    return models, kinds, rate


def _run(g):
    models, kinds, rate = _models(g)
    for m in models:
        m.ShapeFunctions()
        m.AssembleLumpedMass()
    if len(models) > 1:
        po.halo_sum(models, "mass")
    d0 = rank_dict(g, 0)
    n, dth, eh = po.run_explicit(models, kinds, rate, float(g["param_tMax"]), int(d0["steps"][0]))
    return models, n, dth, eh


@pytest.mark.parametrize("name", SINGLE + MULTI)
def test_oracle_bit_exact_vs_reference(name):
    g = golden(name)
    models, n, dth, eh = _run(g)
    for r, m in enumerate(models):
        d = rank_dict(g, r)
        assert n == int(d["steps"][0])
        assert m.Time == float(d["Time"][0]) and m.dt == float(d["dt"][0])
        assert np.array_equal(dth, d["dt_hist"])
        assert np.array_equal(m.mass, d["mass"])
        for k in ["displacements", "velocities", "accelerations", "fi", "f_net", "boundary"]:
            assert np.array_equal(getattr(m, k), d[k]), (name, r, k)
        for k in ["detJacobian", "F", "detF", "pk2", "Hn_1", "Hn_2", "S0n"]:
            if k in d and d[k].size:
                assert np.array_equal(getattr(m, k), d[k]), (name, r, k)
        if "Eavg" in d:
            assert np.array_equal(m.CalculateStrain(), d["Eavg"])
    # energy line written by the reference (%12.6e): time Wint Wext WKE total
    ef = g["energy_file"][-1]
    assert abs(models[0].Time - ef[0]) <= 1e-6 * abs(ef[0])
    for got, want in zip(eh[-1], ef[1:]):
        assert abs(got - want) <= 2e-6 * max(abs(want), 1e-300) + 1e-30


def test_ex9_known_answer_vs_abaqus():
    """examples/ex9/compareResults.py:14-29: last line of plot.dat vs last line
    of abaqus/abaqus.rpt (columns 10-12 = U1..U3 of the node at (L,L,L)):
    time within 1 %, displacements within 5 %."""
    abaqus_t, abaqus_u = 1.0, np.array([-1.03349e-03, 7.0e-03, -1.03349e-03])
    g = golden("ex9_1elt")
    models, n, _, _ = _run(g)
    m = models[0]
    X = m.coordinates.reshape(-1, 3)
    node = int(np.where(np.all(np.abs(X - mesh.CUBE_L) < 1e-5, axis=1))[0][0])
    u = m.displacements.reshape(-1, 3)[node]
    assert abs(m.Time - abaqus_t) * 100.0 / abaqus_t < 1.0
    assert np.max(np.abs((u - abaqus_u) * 100.0 / abaqus_u)) < 5.0
    # value observed when the reference itself was run in this container
    assert np.allclose(u, [-1.07624e-03, 7.00001e-03, -1.07624e-03], rtol=2e-6)


def test_golden_runs_are_nontrivial():
    for name in SINGLE:
        d = rank_dict(golden(name), 0)
        assert np.all(np.isfinite(d["displacements"])) and np.abs(d["displacements"]).max() > 1e-6
        assert int(d["steps"][0]) >= 50


def test_geometry_quirk_signed_parallelogram_test():
    """src/math/Geometry.cpp:46 compares centerD[i] < tol without fabs: a face
    whose centre offset is large and NEGATIVE still takes the parallelogram
    branch.  The oracle must reproduce that, not 'fix' it."""
    L = po.lib()
    c = np.array([[0, 0, 0], [1, 0, 0], [1.5, 1.5, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    idx = np.array([0, 1, 2, 3], dtype=np.int32)
    a_pos = L.oracle_areaHexahedronFace(po._d(c.reshape(-1).copy()), po._i(idx))
    c2 = c.copy()
    c2[2] = [0.5, 0.5, 0]  # centerD negative -> parallelogram formula (quirk)
    a_neg = L.oracle_areaHexahedronFace(po._d(c2.reshape(-1).copy()), po._i(idx))
    p1, p2, p3, p4 = c2[0], c2[1], c2[2], c2[3]
    c1v, c2v = 0.25 * (-p1 + p2 + p3 - p4), 0.25 * (-p1 - p2 + p3 + p4)
    assert a_neg == pytest.approx(4.0 * np.linalg.norm(np.cross(c1v, c2v)), rel=1e-14)
    assert a_pos == pytest.approx(1.5, rel=1e-12)  # planar quad: 2x2 Gauss is exact (shoelace area 1.5)


# ---- injury criteria: ex5.cpp:1311-1430 over the reference's own
#      CalculateMaximumPrincipalStrain / compute95thPercentileValue / computePartVolume ---------------------
INJURY = ["inj6_p1", "inj6_p3", "inj6b_p1"]


def run_injury_oracle(g):
    models, kinds, rate = _models(g)
    for m in models:
        m.ShapeFunctions()
        m.AssembleLumpedMass()
    if len(models) > 1:
        po.halo_sum(models, "mass")
    inj = [po.InjuryCriteria(m, exclude_pids=g["param_exclude"]) for m in models]
    d0 = rank_dict(g, 0)
    n, dth, eh = po.run_explicit(models, kinds, rate, float(g["param_tMax"]), int(d0["steps"][0]), injuries=inj)
    return models, inj, n, dth


@pytest.mark.parametrize("name", INJURY)
def test_oracle_injury_bit_exact_vs_reference(name):
    g = golden(name)
    models, inj, n, dth = run_injury_oracle(g)
    for r, (m, q) in enumerate(zip(models, inj)):
        d = rank_dict(g, r)
        assert n == int(d["steps"][0]) and np.array_equal(dth, d["dt_hist"])
        assert np.array_equal(m.displacements, d["displacements"])
        assert np.array_equal(q.get("elementIDInjury"), d["inj_elems"])
        for ours, ref in (("MPSgt15", "inj_gt15"), ("MPSgt30", "inj_gt30"), ("MPSRgt120", "inj_r120"),
                          ("MPSxSRgt28", "inj_xsr28"), ("PS_Old", "inj_ps_old"), ("PSxSRArray", "inj_psxsr")):
            assert np.array_equal(q.get(ours), d[ref]), (name, r, ours)
        assert np.array_equal(q.scalars(), d["inj_scalars"]), (name, r)
        assert np.array_equal(q.extreme_elems(), d["inj_extreme_elems"])
        l95, lx95 = q.lists()
        assert np.array_equal(l95, d["inj_list95"]) and np.array_equal(lx95, d["inj_listx95"])
        assert np.array_equal(q.volumes(), d["inj_volumes"])
        if "Eavg" in d:
            for e in (0, 7, m.nElements - 1):
                _, E = q.principal(e)
                assert np.array_equal(E, d["Eavg"][9 * e:9 * e + 9])
    # the fixtures exercise every branch: some but not all flags set, both lists non-empty
    d = rank_dict(g, 0)
    assert 0 < d["inj_xsr28"].sum() < d["inj_xsr28"].size and len(d["inj_list95"]) and len(d["inj_listx95"])


def test_percentile_is_the_order_statistic_the_reference_picks():
    # math.cpp:160-199 (and :235-332, which always falls through to it): element (int)(0.95 n) - 1, ascending
    rng = np.random.default_rng(5)
    a, b = rng.random(1011), rng.random(1031) * 2.0
    want = np.sort(np.concatenate([a, b]))[int((1011 + 1031) * 0.95) - 1]
    assert po.percentile95([a, b]) == want
    g = golden("inj6_p3")
    h = rank_dict(g, 0)["inj_hist95"]
    assert all(np.array_equal(rank_dict(g, r)["inj_hist95"], h) for r in range(3))  # collective: same on all ranks


def test_principal_strain_degenerate_states():
    # CalculateStrain.cpp:36-41: exactly diagonal E takes the shortcut; a double root that is not exactly diagonal can
    # push acos's argument past 1 -> NaN -> the reference reports 0 (fmax/fmin drop NaN only if one operand is finite).
    X, conn, pid = mesh.cube_mesh(1)
    m = po.OracleModel(X, conn, pid, [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])
    m.ShapeFunctions()
    q = po.InjuryCriteria(m)
    Fd = np.diag([1.2, 0.9, 0.9])
    m.F[:] = np.tile(Fd.T.reshape(-1), 8)
    (smax, smin, shear), E = q.principal(0)
    assert smax == pytest.approx(0.5 * (1.44 - 1), rel=1e-14) and smin == pytest.approx(0.5 * (0.81 - 1), rel=1e-14)
    assert shear == pytest.approx(0.5 * (smax - smin), rel=1e-14)


# ---- rigid-body prescribed motion: ex5.cpp:339-371,976-1020 --------------------------------
def rigid_tables(g):
    return [(g["param_rigid_t%d" % k], g["param_rigid_v%d" % k]) for k in range(6)]


def test_oracle_rigid_body_bc_bit_exact_vs_reference_library():
    """The ex5 loop with ApplyAccBoundaryConditions: the fixture was written by the reference library (GetForce, energy,
    quaternion and interpolation helpers, injury functions) under the harness driver; the oracle reproduces it bit for
    bit.  The Dormand-Prince step itself is the same restatement on both sides (Boost is absent): unpinned."""
    g = golden("rigid6_p1")
    d = rank_dict(g, 0)
    m = po.OracleModel(d["coordinates"], d["connectivity"], d["pid"], d["materialID"], d["properties"])
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    rb = po.RigidBody(m, rigid_tables(g))
    inj = po.InjuryCriteria(m, exclude_pids=g["param_exclude"])
    n, dth, eh = po.run_explicit_rigid(m, rb, float(g["param_tMax"]), int(d["steps"][0]), failure_dt=1e-11, injury=inj)
    assert n == int(d["steps"][0]) and np.array_equal(dth, d["dt_hist"])
    assert np.array_equal(rb.boundaryID, d["rb_boundaryID"])
    assert np.array_equal(rb.y, d["rb_y"]) and np.array_equal(rb.ydot, d["rb_ydot"])
    for k in ["displacements", "velocities", "accelerations", "fi", "boundary", "pk2"]:
        assert np.array_equal(getattr(m, k), d[k]), k
    assert np.array_equal(inj.get("PS_Old"), d["inj_ps_old"]) and np.array_equal(inj.scalars(), d["inj_scalars"])
    ef = g["energy_file"][-1]
    for got, want in zip(eh[-1], ef[1:]):
        assert abs(got - want) <= 2e-6 * max(abs(want), 1e-300) + 1e-30


def test_dopri5_restatement_is_fifth_order_and_fsal():
    """No Boost here, so the integrator is checked against what it must satisfy: with constant angular acceleration about
    a fixed axis the states are polynomials of degree <= 2 (exact for a 5th-order method), the rotation generator is
    0.5 * angle * axis (math.cpp:122-132 uses the half-angle convention), and ydot out is f(y_new, t + dt)."""
    X, conn, pid = mesh.cube_mesh(1)
    m = po.OracleModel(X, conn, pid, [0], [1000.0, 0, 0, 0, 0, 0, 0, 0, 0])
    m.ShapeFunctions()
    al, acc = 300.0, 20.0
    tabs = [([0.0, 1.0], [0.0, 0.0]), ([0.0, 1.0], [0.0, 0.0]), ([0.0, 1.0], [al, al]),
            ([0.0, 1.0], [acc, acc]), ([0.0, 1.0], [0.0, 0.0]), ([0.0, 1.0], [0.0, 0.0])]
    rb = po.RigidBody(m, tabs)
    # the driver starts from ydot = 0 (ex5.cpp:897-900), i.e. the first step uses k1 = 0, not f(y0, 0): start one
    # step later with a consistent FSAL derivative instead
    t, dt = 0.0, 1e-3
    rb.step(t, dt)
    t += dt
    y1 = rb.y
    for _ in range(50):
        rb.step(t, dt)
        t += dt
    y = rb.y
    T = t - dt  # time since the consistent start
    w1, th1 = y1[2], 2.0 * y1[5]
    assert y[2] == pytest.approx(w1 + al * T, rel=1e-13)                       # omega_z
    assert 2.0 * y[5] == pytest.approx(th1 + w1 * T + 0.5 * al * T * T, rel=1e-12)  # angle = 2 |r|
    assert y[6] == pytest.approx(y1[6] + acc * T, rel=1e-13)
    assert y[9] == pytest.approx(y1[9] + y1[6] * T + 0.5 * acc * T * T, rel=1e-12)
    assert rb.ydot[2] == al and rb.ydot[9] == y[6]


# ---- mixed C3D8 / C3D4 meshes -------------------------------------------------------------------
MIXED = ["mix4_p1", "mix4v_p1"]


def mixed_model(d):
    return po.OracleModel(d["coordinates"], d["connectivity"], d["pid"], d["materialID"], d["properties"], eptr=d["eptr"])


@pytest.mark.parametrize("name", MIXED)
def test_oracle_mixed_hex_tet_bit_exact_vs_reference(name):
    """One-point tetrahedra (ShapeFunction_C3D4.cpp, CalculateCharacteristicLength_C3D4.cpp, GaussQuadrature3D.cpp:62-69)
    next to hexahedra: mass, dt history, end state, per-Gauss-point arrays in the reference's packed layouts, Prony
    history, and the injury criteria with one Gauss point per element."""
    g = golden(name)
    d = rank_dict(g, 0)
    m = mixed_model(d)
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    assert np.array_equal(m.mass, d["mass"]) and np.array_equal(m.detJacobian, d["detJacobian"])
    kind, rate = mesh.benchmark_bc(d["coordinates"], dMax=float(g["param_dMax"]), tMax=float(g["param_tMax"]))
    inj = po.InjuryCriteria(m, exclude_pids=[])
    n, dth, eh = po.run_explicit([m], [kind], rate, float(g["param_tMax"]), int(d["steps"][0]), injuries=[inj])
    assert n == int(d["steps"][0]) and np.array_equal(dth, d["dt_hist"])
    for k in ["displacements", "velocities", "accelerations", "fi", "f_net", "F", "detF", "pk2", "Hn_1", "Hn_2", "S0n"]:
        if k in d and d[k].size:
            assert np.array_equal(getattr(m, k), d[k]), (name, k)
    assert np.array_equal(m.CalculateStrain(), d["Eavg"])
    assert np.array_equal(inj.get("PS_Old"), d["inj_ps_old"]) and np.array_equal(inj.scalars(), d["inj_scalars"])
    assert np.array_equal(inj.volumes(), d["inj_volumes"])
    assert (np.diff(d["eptr"]) == 4).sum() > 0 and (np.diff(d["eptr"]) == 8).sum() > 0
