"""Step ring (ftb200_step_ring): the device writes every finished step's scalars into pinned host memory; the records
must be the very numbers the device-side histories and the blocking poll report, in order, with wrap-around."""
import numpy as np
import pytest

from femtech_b200 import mesh

pytestmark = pytest.mark.gpu

SOFT = [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0]


def model(n=6, jitter=0.05):
    from femtech_b200 import solver
    X, conn, pid = mesh.cube_mesh(n, jitter=jitter)
    m = solver.FemTech(X, conn, pid, [1], SOFT)
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    kind, rate = mesh.benchmark_bc(X, dMax=0.007, tMax=0.004)
    m.set_bc(kind, rate)
    return m


@pytest.mark.parametrize("energy", [1, 0])
def test_ring_records_equal_histories(energy):
    nsteps = 70  # spans three 25-step graph replays
    m = model()
    ring = m.step_ring(128)
    m.explicit_begin(energy_every=energy, record_steps=nsteps)
    m.run_async(1.0, nsteps)
    recs = np.array([m.wait_step(k) for k in range(1, nsteps + 1)])
    m._poll()
    assert m.steps_done == nsteps
    dth, eh = m.history(0, nsteps)
    assert np.array_equal(recs[:, 2], np.arange(1, nsteps + 1))
    assert np.all(recs[:, 3] == 0)
    # record k holds the dt of step k + 1 (the value explicit_poll returns as dt after step k)
    assert np.array_equal(recs[:-1, 1], dth[1:])
    assert recs[-1, 1] == m.dt and recs[-1, 0] == m.Time
    assert np.allclose(np.diff(recs[:, 0]), dth[1:], rtol=1e-12, atol=0)
    if energy:
        assert np.array_equal(recs[:, 4:8], eh.reshape(-1, 4))
        assert np.array_equal(recs[-1, 4:8], m.energy())
    assert np.all(ring[nsteps:, 2] == -1.0)  # untouched slots keep the empty marker
    m.close()


def test_ring_wraps_and_survives_a_second_run():
    m = model(n=4)
    ring = m.step_ring(16)
    m.explicit_begin(energy_every=1)
    m.run_async(1.0, 40)
    last = m.wait_step(40)
    m._poll()
    assert last[2] == 40 and last[0] == m.Time
    # the ring holds the last 16 steps, each in slot (k - 1) % 16
    for k in range(25, 41):
        assert ring[(k - 1) % 16, 2] == k
    with pytest.raises(Exception):
        m.wait_step(3, timeout_s=0.5)  # overwritten long ago
    m.run_async(1.0, 5)  # a second call continues the numbering
    assert m.wait_step(45)[2] == 45
    m.step_ring(0)
    m.run_async(1.0, 3)  # released: the loop runs on without a ring
    m._poll()
    assert m.steps_done == 48
    m.close()
