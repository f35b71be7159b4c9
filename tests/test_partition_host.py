"""Host-side multi-rank logic (no GPU): partition maps bit-exact against the reference's
PartitionMesh output (golden vectors from 2/3/4/8-rank reference runs through the MPI shim),
the analytic brick split against the general map builder, and a world_size-2 gloo run of
the halo exchange checked against the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, rank_dict
from femtech_b200 import dist as fdist
from femtech_b200 import mesh


def _global_conn_from_golden(g):
    P = int(g["nranks"])
    nE = sum(rank_dict(g, r)["global_eid"].size for r in range(P))
    gconn = np.zeros((nE, 8), dtype=np.int64)
    for r in range(P):
        d = rank_dict(g, r)
        gn = d["globalNodeID"].astype(np.int64) - 1
        gconn[d["global_eid"] - 1] = gn[d["connectivity"].reshape(-1, 8)]
    return gconn


@pytest.mark.parametrize("name", ["bench10_p2", "bench10_p4", "bench10_p8", "cube6mix_p3"])
def test_node_maps_bit_exact_vs_reference(name):
    """Given the reference's element distribution (ParMETIS part + migration order, an input here), the
    local numbering and every send list must equal PartitionMesh's output bit for bit."""
    g = golden(name)
    P = int(g["nranks"])
    gconn = _global_conn_from_golden(g)
    maps = fdist.maps_from_elements(gconn, [rank_dict(g, r)["global_eid"] - 1 for r in range(P)])
    for r in range(P):
        d = rank_dict(g, r)
        assert np.array_equal(maps[r]["connectivity"], d["connectivity"])
        assert np.array_equal(maps[r]["globalNodeID"], d["globalNodeID"])
        assert np.array_equal(maps[r]["sendProcessID"], d["sendProcessID"])
        assert np.array_equal(maps[r]["sendNeighbourCountCum"], d["sendNeighbourCountCum"])
        assert np.array_equal(maps[r]["sendNodeIndex"], d["sendNodeIndex"])


@pytest.mark.parametrize("P", [2, 4, 8])
def test_brick_partition_matches_general_builder(P):
    pg = fdist.proc_grid(P)
    parts = [fdist.brick_partition(3, pg, r) for r in range(P)]
    # assemble the global mesh from the bricks and rebuild the maps with the reference's rules
    gid_all = np.unique(np.concatenate([p["node_gids"] for p in parts]))
    gconn, owners = [], []
    for r, p in enumerate(parts):
        gconn.append(p["node_gids"][p["connectivity"]])
        owners.append(np.arange(len(gconn[-1])) + sum(len(x) for x in gconn[:-1]))
    gconn = np.concatenate(gconn)
    compact = np.searchsorted(gid_all, gconn)
    maps = fdist.maps_from_elements(compact, owners)
    for r, p in enumerate(parts):
        assert np.array_equal(maps[r]["connectivity"].reshape(-1, 8), p["connectivity"])
        for k in ("sendProcessID", "sendNeighbourCountCum", "sendNodeIndex"):
            assert np.array_equal(maps[r][k], p["comm"][k]), (r, k)
    assert fdist.proc_grid(8) == (2, 2, 2) and fdist.proc_grid(4) == (2, 2, 1) and fdist.proc_grid(2) == (2, 1, 1)


def _oracle_models(parts, matid, props):
    from oracle import pyoracle as po
    ms = []
    for r, p in enumerate(parts):
        m = po.OracleModel(p["coordinates"], p["connectivity"], p["pid"], matid, props, comm=p["comm"], world_rank=r)
        m.ShapeFunctions()
        m.AssembleLumpedMass()
        ms.append(m)
    return ms


def test_multirank_emulation_agrees_with_single_rank():
    """Oracle on 4 bricks (halo sums emulated) vs oracle on the assembled global mesh."""
    from oracle import pyoracle as po
    P, n = 4, 3
    pg = fdist.proc_grid(P)
    parts = [fdist.brick_partition(n, pg, r) for r in range(P)]
    soft = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]
    ms = _oracle_models(parts, [1], soft)
    po.halo_sum(ms, "mass")
    Ly = parts[0]["box"][1]
    kinds = []
    for p in parts:
        k, rate = mesh.benchmark_bc(p["coordinates"], L=Ly)
        kinds.append(k)
    n_steps, _, _ = po.run_explicit(ms, kinds, rate, 0.1, 40)
    # global mesh
    gid_all = np.unique(np.concatenate([p["node_gids"] for p in parts]))
    Xg = np.zeros((gid_all.size, 3))
    gconn = []
    for p in parts:
        loc = np.searchsorted(gid_all, p["node_gids"])
        Xg[loc] = p["coordinates"]
        gconn.append(loc[p["connectivity"]])
    gconn = np.concatenate(gconn)
    og = po.OracleModel(Xg, gconn, np.zeros(len(gconn), np.int32), [1], soft)
    og.ShapeFunctions()
    og.AssembleLumpedMass()
    kg, rate = mesh.benchmark_bc(Xg, L=Ly)
    ng, _, _ = po.run_explicit([og], [kg], rate, 0.1, 40)
    assert ng == n_steps == 40
    for p, m in zip(parts, ms):
        loc = np.searchsorted(gid_all, p["node_gids"])
        ug = og.displacements.reshape(-1, 3)[loc]
        assert np.abs(m.displacements.reshape(-1, 3) - ug).max() < 1e-12 * np.abs(og.displacements).max()


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = fdist.proc_grid(world)
    part = fdist.brick_partition(3, pg, rank)
    soft = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]
    m = po.OracleModel(part["coordinates"], part["connectivity"], part["pid"], [1], soft, comm=part["comm"],
                       world_rank=rank)
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    halo = fdist.HaloExchange(part["comm"], torch.device("cpu"), dist)
    idx = part["comm"]["sendNodeIndex"]

    def halo_sum(field):
        halo.send[:3 * idx.size] = torch.from_numpy(field.reshape(-1, 3)[idx].reshape(-1).copy())
        halo.exchange()
        fdist.halo_add_host(field, part["comm"], halo.recv[:3 * idx.size].numpy())

    halo_sum(m.mass)
    kind, rate = mesh.benchmark_bc(part["coordinates"], L=part["box"][1])
    bc = kind > 0
    # three explicit steps by hand: oracle for the local work, gloo for the exchange and the dt MIN
    m.boundary[bc] = 1
    m.velocities[bc] = rate[kind[bc]]
    Time = 0.0

    def global_dt():
        t = torch.tensor([m.StableTimeStep()], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return 0.8 * float(t[0])

    def get_force():
        m.L.oracle_GetForce_local(po.C.byref(m.s))
        halo_sum(m.fi)
        m.L.oracle_GetForce_finish(po.C.byref(m.s))

    dt = global_dt()
    m.s.dt = dt
    get_force()
    m.CalculateAccelerations()
    for _ in range(3):
        t_n, t_np1 = Time, Time + dt
        Time = t_np1
        t_half = 0.5 * (t_np1 + t_n)
        free = m.boundary == 0
        vh = np.where(free, m.velocities + (t_half - t_n) * m.accelerations, m.velocities)
        m.displacements[free] = m.displacements[free] + dt * vh[free]
        m.displacements[bc] = Time * rate[kind[bc]]
        get_force()
        m.CalculateAccelerations()
        m.velocities[free] = vh[free] + (t_np1 - t_half) * m.accelerations[free]
        dt = global_dt()
        m.s.dt = dt
    q.put((rank, m.displacements.copy(), m.velocities.copy(), m.mass.copy(), dt))
    dist.destroy_process_group()


def test_gloo_world2_halo_exchange_matches_oracle_emulation():
    """world_size 2 over gloo on CPU: HaloExchange + dt MIN reproduce the in-process emulation bit for bit."""
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, u, v, mass, dt = q.get(timeout=120)
        res[r] = (u, v, mass, dt)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [fdist.brick_partition(3, fdist.proc_grid(world), r) for r in range(world)]
    ms = _oracle_models(parts, [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])
    po.halo_sum(ms, "mass")
    kinds = [mesh.benchmark_bc(p["coordinates"], L=p["box"][1])[0] for p in parts]
    rate = mesh.benchmark_bc(parts[0]["coordinates"], L=parts[0]["box"][1])[1]
    n, _, _ = po.run_explicit(ms, kinds, rate, 1.0, 3)
    assert n == 3
    for r in range(world):
        u, v, mass, dt = res[r]
        assert np.array_equal(mass, ms[r].mass)
        assert np.array_equal(u, ms[r].displacements)
        assert np.array_equal(v, ms[r].velocities)
        assert dt == ms[r].dt
