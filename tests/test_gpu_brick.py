"""GPU parity of the brick-fused step (k_brick + k_surf, femtech_b200/csrc/ftb200_brick.cuh).

The brick path keeps the element forces on the SM: a thread block integrates one brick of elements, assembles and finishes
the nodes interior to the brick and leaves one partial sum per surface node for a thin second pass.  Checked here through
the C-ABI: (1) against the CPU oracle on the same inputs at the north-star tolerance (u, v, PK2 <= 1e-9), (2) against the
two-kernel step on the same mesh (FTB200_BRICK=0) at rounding level -- the only arithmetic difference is the summation
order at surface nodes --, (3) chunked runs and single-step calls against one long run (bit-identical: the state between
steps is the full-step state), (4) the decomposition itself (every node finished exactly once), (5) ragged and tiny bricks
(FTB200_BRICK_DIMS) and sheared parallelepiped meshes, (6) meshes that do not qualify fall back to the two-kernel step."""
import os

import numpy as np
import pytest

from femtech_b200 import mesh

pytestmark = pytest.mark.gpu

TOL = 1e-9
SOFT = [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0]
BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]
HGO = BRAIN[:4] + [10.0, 0, 0, 0, 0]
PROPS = {1: SOFT, 4: HGO, 5: BRAIN}


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.kw.setdefault("FTB200_BRICK", "1")
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make(X, conn, pid, matid, props, kind, rate, nsteps, energy=1):
    from femtech_b200 import solver
    m = solver.FemTech(X, conn, pid, matid, props)
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    m.set_bc(kind, rate)
    m.explicit_begin(energy_every=energy, record_steps=nsteps)
    return m


def run_oracle(X, conn, pid, matid, props, kind, rate, nsteps):
    from oracle import pyoracle as po
    o = po.OracleModel(X, conn, pid, matid, props)
    o.ShapeFunctions()
    o.AssembleLumpedMass()
    n, dth, _ = po.run_explicit([o], [kind], rate, 1.0, nsteps)
    assert n == nsteps
    return o, dth


@pytest.mark.parametrize("mat,n,dims", [(1, 8, "10,5,5"), (4, 8, "10,5,5"), (1, 7, "3,2,2"), (4, 6, "2,3,1")])
def test_brick_step_matches_oracle(mat, n, dims):
    X, conn, pid = mesh.cube_mesh(n)
    kind, rate = mesh.benchmark_bc(X, dMax=0.007, tMax=0.1 if mat == 1 else 0.004)
    nsteps = 150
    o, dth_o = run_oracle(X, conn, pid, [mat], PROPS[mat], kind, rate, nsteps)
    with env(FTB200_BRICK_DIMS=dims):
        m = make(X, conn, pid, [mat], PROPS[mat], kind, rate, nsteps)
    info = m.brick_info
    assert info["active"] and info["bricks"] >= 2
    assert info["interior_nodes"] + info["surface_nodes"] == X.shape[0]
    assert m.ExplicitDynamics(1.0, maxSteps=nsteps) == nsteps
    dth, eh = m.history(0, nsteps)
    assert np.allclose(dth, dth_o, rtol=1e-11, atol=0)
    assert rel(m.displacements, o.displacements) < TOL
    assert rel(m.velocities, o.velocities) < TOL
    assert rel(m.accelerations, o.accelerations) < 1e-6
    out = m.gp_outputs()
    assert rel(out["pk2"], o.pk2) < TOL
    assert rel(out["F"], o.F) < TOL
    m.close()


@pytest.mark.parametrize("mat,energy", [(1, 1), (1, 0), (4, 1)])
def test_brick_step_agrees_with_two_kernel_step(mat, energy):
    """Same mesh, same inputs: brick-fused step against k_elem_affine + k_node.  12^3 with 10 x 5 x 5 bricks has full,
    ragged and thin bricks.  State, dt history, energies and the lazily rebuilt internal force agree to rounding."""
    X, conn, pid = mesh.cube_mesh(12)
    kind, rate = mesh.benchmark_bc(X, dMax=0.02, tMax=0.004)
    nsteps = 100
    res = []
    for flag in ("1", "0"):
        with env(FTB200_BRICK=flag):
            m = make(X, conn, pid, [mat], PROPS[mat], kind, rate, nsteps, energy=energy)
        assert m.brick_info["active"] == (flag == "1")
        assert m.ExplicitDynamics(1.0, maxSteps=nsteps) == nsteps
        dth, eh = m.history(0, nsteps)
        fi, fnet = m.fi.copy(), m.f_net.copy()  # rebuilt lazily by sync_out (brick mode: one more force evaluation)
        res.append({"u": m.displacements.copy(), "v": m.velocities.copy(), "a": m.accelerations.copy(), "t": m.Time,
                    "dt": dth.copy(), "e": eh.copy(), "fi": fi, "fnet": fnet, "pk2": m.gp_outputs()["pk2"].copy()})
        m.close()
    b, g = res
    assert abs(b["t"] - g["t"]) <= 1e-13 * abs(g["t"])
    assert np.allclose(b["dt"], g["dt"], rtol=1e-12, atol=0)
    for k in ("u", "v", "pk2"):
        assert rel(b[k], g[k]) < 1e-11, k
    for k in ("a", "fi", "fnet"):
        assert rel(b[k], g[k]) < 1e-8, k
    if energy:
        assert np.allclose(b["e"], g["e"], rtol=1e-9, atol=1e-12 * np.abs(g["e"]).max())


def test_brick_chunked_runs_equal_one_run():
    """25-step graphs, single steps and a ragged tail: the same bits as one 60-step call."""
    X, conn, pid = mesh.cube_mesh(10)
    kind, rate = mesh.benchmark_bc(X, dMax=0.02, tMax=0.004)
    with env():
        ref = make(X, conn, pid, [1], SOFT, kind, rate, 60)
        m = make(X, conn, pid, [1], SOFT, kind, rate, 60)
    assert ref.brick_info["active"]
    assert ref.ExplicitDynamics(1.0, maxSteps=60) == 60
    done = 0
    for chunk in (1, 1, 26, 7, 25):
        done += m.ExplicitDynamics(1.0, maxSteps=chunk)
    assert done == 60
    assert np.array_equal(m.displacements, ref.displacements)
    assert np.array_equal(m.velocities, ref.velocities)
    assert np.array_equal(m.accelerations, ref.accelerations)
    assert m.Time == ref.Time
    d0, e0 = ref.history(0, 60)
    d1, e1 = m.history(0, 60)
    assert np.array_equal(d0, d1) and np.array_equal(e0, e1)
    ref.close(); m.close()


def test_brick_decomposition_covers_every_element_and_node_once():
    X, conn, pid = mesh.cube_mesh(12)
    kind, rate = mesh.benchmark_bc(X, dMax=0.02, tMax=0.004)
    with env():
        m = make(X, conn, pid, [1], SOFT, kind, rate, 1)
    info = m.brick_info
    be, bn = m.brick_maps()
    assert be.min() == 0 and be.max() == info["bricks"] - 1
    assert np.bincount(be).max() <= 256
    # a node is interior to brick b exactly when all of its elements lie in b
    nb = [set() for _ in range(X.shape[0])]
    for e, row in enumerate(conn):
        for nd in row:
            nb[nd].add(int(be[e]))
    for n, s in enumerate(nb):
        assert bn[n] == (next(iter(s)) if len(s) == 1 else -1)
    assert info["interior_nodes"] == int((bn >= 0).sum())
    assert info["partial_slots"] == sum(len(s) for s in nb if len(s) > 1)
    m.close()


def test_sheared_parallelepiped_mesh_takes_the_brick_path():
    """Affine image of the cube on a dyadic grid (edge vectors stay bit-equal): every element is a parallelepiped but not a box."""
    X, conn, pid = mesh.cube_mesh(8)
    h = X[:, 0].max() / 8
    G = np.rint(X / h)  # integer lattice
    D = np.array([[64, 8, 0], [-4, 64, 12], [16, 0, 64]]) / 1024.0
    Xs = G @ D
    kind, rate = mesh.benchmark_bc(X, dMax=0.007, tMax=0.1)
    nsteps = 80
    o, dth_o = run_oracle(Xs, conn, pid, [1], SOFT, kind, rate, nsteps)
    with env(FTB200_BRICK_DIMS="4,4,4"):
        m = make(Xs, conn, pid, [1], SOFT, kind, rate, nsteps)
    assert m.affine_elements == conn.shape[0] and m.brick_info["active"]
    assert m.ExplicitDynamics(1.0, maxSteps=nsteps) == nsteps
    assert rel(m.displacements, o.displacements) < TOL
    assert rel(m.velocities, o.velocities) < TOL
    assert rel(m.gp_outputs()["pk2"], o.pk2) < TOL
    m.close()


def test_meshes_that_do_not_qualify_keep_the_two_kernel_step():
    Xj, conn, pid = mesh.cube_mesh(6, jitter=0.05)
    kind, rate = mesh.benchmark_bc(Xj, dMax=0.007, tMax=0.004)
    with env():
        m = make(Xj, conn, pid, [1], SOFT, kind, rate, 1)
    assert not m.brick_info["active"] and m.brick_info["bricks"] == 0
    m.close()
    X, conn, pid = mesh.cube_mesh(6)
    with env():
        m = make(X, conn, pid, [5], BRAIN, kind, rate, 1)
    assert not m.brick_info["active"]
    m.close()
