"""Overlapped step (FTB200_OVERLAP=1: k_node_ovl beside the element kernel, START half behind k_adv) against the serial
step on the same inputs.  Same arithmetic, same summation order, same energy partials: the states, the dt history and the
energies must be IDENTICAL bit for bit, for both element kernels, with and without the energy check, across graph
replays, chunk boundaries (several 16384-element chunks) and the end of a run in the middle of a graph."""
import os

import numpy as np
import pytest

from femtech_b200 import mesh

pytestmark = pytest.mark.gpu

SOFT = [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0]
BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]
HGO = BRAIN[:4] + [10.0, 0, 0, 0, 0]


def run(X, conn, pid, mat, props, nsteps, energy, env):
    from femtech_b200 import solver
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        m = solver.FemTech(X, conn, pid, [mat], props)
        m.ShapeFunctions()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    m.AssembleLumpedMass()
    kind, rate = mesh.benchmark_bc(X, dMax=0.02, tMax=0.004)
    m.set_bc(kind, rate)
    m.explicit_begin(energy_every=energy, record_steps=nsteps)
    done = 0
    for chunk in (nsteps - 33, 33):  # two calls: the second ends in the middle of a 25-step graph
        done += m.ExplicitDynamics(1.0, maxSteps=chunk)
    assert done == nsteps
    dth, eh = m.history(0, nsteps)
    out = dict(u=m.displacements.copy(), v=m.velocities.copy(), a=m.accelerations.copy(), fi=m.fi.copy(), dth=dth, eh=eh,
               T=m.Time, launches=m.gpu_launches, status=m.status_bits)
    m.close()
    return out


@pytest.mark.parametrize("mat,jitter,energy,extra", [(1, 0.0, 1, {}), (1, 0.05, 1, {}), (4, 0.0, 0, {}),
                                                     (1, 0.0, 1, {"FTB200_OVL_NODE_BLOCKS": "2", "FTB200_OVL_ELEM_BLOCKS": "6"})])
def test_overlapped_step_is_bit_identical_to_serial_step(mat, jitter, energy, extra):
    X, conn, pid = mesh.cube_mesh(40, jitter=jitter)  # 64000 elements: 4 chunks, 539 node tiles
    props = SOFT if mat == 1 else HGO
    nsteps = 83
    ser = run(X, conn, pid, mat, props, nsteps, energy, {"FTB200_OVERLAP": "0"})
    env = {"FTB200_OVERLAP": "1"}
    env.update(extra)
    ovl = run(X, conn, pid, mat, props, nsteps, energy, env)
    assert ovl["status"] == 0 and ser["status"] == 0
    assert ovl["launches"] > ser["launches"]  # five kernels per step instead of four: the overlapped path really ran
    assert ovl["T"] == ser["T"]
    for k in ("dth", "u", "v", "a", "fi"):
        assert np.array_equal(ovl[k], ser[k]), k
    if energy:
        assert np.array_equal(ovl["eh"], ser["eh"])


def test_overlap_request_is_ignored_where_it_does_not_apply():
    """Two materials = two runs of elements: the request falls back to the serial step (same results, same launches)."""
    X, conn, pid = mesh.cube_mesh(12, nparts_z=2)
    from femtech_b200 import solver
    res = []
    for flag in ("0", "1"):
        os.environ["FTB200_OVERLAP"] = flag
        try:
            m = solver.FemTech(X, conn, pid, [1, 4], SOFT + HGO)
            m.ShapeFunctions()
        finally:
            del os.environ["FTB200_OVERLAP"]
        m.AssembleLumpedMass()
        kind, rate = mesh.benchmark_bc(X, dMax=0.02, tMax=0.004)
        m.set_bc(kind, rate)
        m.explicit_begin(energy_every=1)
        m.ExplicitDynamics(1.0, maxSteps=30)
        res.append((m.displacements.copy(), m.gpu_launches))
        m.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
