"""The reference's UNMODIFIED example drivers, linked against femtech_b200 (integration/femtech_host.cpp ->
C-ABI -> CUDA), run on the GPU: examples/ex9/ex9.cpp is the reference's own known-answer test
(.travis.yml:113-115, compareResults.py) and examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp is
BASELINE config 1.  The binaries are linked in the build container (integration/build_dropin.sh) and travel
to the GPU box under oracle/_ref/."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from femtech_b200 import mesh

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref")


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip("%s not built (integration/build_dropin.sh needs the reference sources)" % name)
    return p


def test_ex9_unmodified_driver_on_gpu(tmp_path):
    exe = _need("dropin_ex9")
    X, conn, pid = mesh.cube_mesh(1)
    mesh.write_abaqus_inp(str(tmp_path / "cube1.inp"), X, conn, pid)
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])
    os.makedirs(tmp_path / "results" / "vtu")  # WriteVTU.cpp:21-30 fopens ./results/vtu/... without creating it
    r = subprocess.run([exe, "cube1.inp"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    last = open(tmp_path / "plot.dat").read().strip().splitlines()[-1].split()
    t, u = float(last[0]), np.array([float(x) for x in last[1:4]])
    # compareResults.py:14-29 against abaqus/abaqus.rpt (5 %), and the value the reference itself produces
    abaqus = np.array([-1.03349e-03, 7.0e-03, -1.03349e-03])
    assert abs(t - 1.0) < 0.01 and np.max(np.abs((u - abaqus) / abaqus)) < 0.05
    assert np.allclose(u, [-1.07624e-03, 7.00001e-03, -1.07624e-03], rtol=2e-5)


def test_benchmarking_parallel_unmodified_driver_on_gpu(tmp_path):
    exe = _need("dropin_benchmarking_parallel")
    from oracle import pyoracle as po
    X, conn, pid = mesh.cube_mesh(10)
    soft = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]
    mesh.write_abaqus_inp(str(tmp_path / "cube10.inp"), X, conn, pid)
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), [1], soft)
    r = subprocess.run([exe, "cube10.inp"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    log = [f for f in os.listdir(tmp_path) if f.startswith("femtech_")][0]
    txt = open(tmp_path / log).read()
    m = re.findall(r"Total time steps : (\d+)", txt)
    vals = re.findall(r"INFO\s*\[\s*0\]:\s+([-0-9.e+]+)\s+([-0-9.e+]+)\s+([-0-9.e+]+)\s+([-0-9.e+]+)\s*$", txt, flags=re.M)
    assert m and vals
    o = po.OracleModel(X, conn, pid, [1], soft)
    o.ShapeFunctions()
    o.AssembleLumpedMass()
    kind, rate = mesh.benchmark_bc(X)
    n, _, eh = po.run_explicit([o], [kind], rate, 0.1, 10 ** 6)
    assert int(m[-1]) == n + 1  # the driver's counter starts at 1
    got = np.array([float(x) for x in vals[-1]])
    want = np.array([o.Time, o.displacements[0], o.displacements[1], o.displacements[2]])
    assert np.allclose(got, want, rtol=2e-3, atol=1e-9)  # the driver logs %11.3e
    en = [l.split() for l in open(tmp_path / [f for f in os.listdir(tmp_path) if f.startswith("energy_")][0])
          if not l.startswith("#")]
    # CheckEnergy writes a line whenever step % nsteps_plot == 0 (Benchmarking-Parallel.cpp:154-155)
    nsteps_plot = int(int(0.1 / dth0(o, X, kind, rate)) / 50)
    last = np.array(en[-1], dtype=float)
    k = (n // nsteps_plot) * nsteps_plot  # 1-based step of the last written line
    assert np.allclose(last[1:4], eh[k - 1][:3], rtol=5e-6, atol=1e-30)


def dth0(o, X, kind, rate):
    """initial dt of the same problem (fresh oracle model)"""
    from oracle import pyoracle as po
    m = po.OracleModel(o.coordinates, o.connectivity, o.pid, o.materialID, o.properties)
    m.ShapeFunctions()
    m.boundary[kind > 0] = 1
    return 0.8 * m.StableTimeStep()


def test_legacy_injury_loop_through_reference_symbols(tmp_path):
    """The oracle harness driver (oracle/ref/ref_dump.cpp: the drivers' time loop + ex5's injury loop, calling
    CalculateMaximumPrincipalStrain / compute95thPercentileValue / computePartVolume by their reference names),
    linked against femtech_b200 instead of the reference's hot path, vs the fixture the all-reference build wrote."""
    exe = _need("dropin_ref_dump")
    from conftest import golden, rank_dict
    from oracle import pyoracle as po
    g = golden("inj6_p1")
    d = rank_dict(g, 0)
    mesh.write_abaqus_inp(str(tmp_path / "cube6.inp"), d["coordinates"].reshape(-1, 3), d["connectivity"].reshape(-1, 8), d["pid"])
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), d["materialID"], d["properties"])
    env = dict(os.environ, REF_INJURY_EXCLUDE=",".join(str(int(p)) for p in g["param_exclude"]))
    r = subprocess.run([exe, "cube6.inp", "out", "400", repr(float(g["param_tMax"])), repr(float(g["param_dMax"]))],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    o = po.read_ref_dump(str(tmp_path / "out.rank0.bin"))
    assert int(o["steps"][0]) == int(d["steps"][0])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(o["displacements"], d["displacements"]) < 1e-9
    assert rel(o["inj_ps_old"], d["inj_ps_old"]) < 1e-9 and rel(o["inj_psxsr"], d["inj_psxsr"]) < 1e-6
    for k in ("inj_elems", "inj_gt15", "inj_gt30", "inj_r120", "inj_xsr28", "inj_list95", "inj_listx95", "inj_extreme_elems"):
        assert np.array_equal(o[k], d[k]), k
    assert rel(o["inj_hist95"], d["inj_hist95"]) < 1e-9 and rel(o["inj_volumes"], d["inj_volumes"]) < 1e-12
    assert rel(o["Eavg"], d["Eavg"]) < 1e-9


def test_mixed_mesh_through_reference_reader_and_symbols(tmp_path):
    """A C3D8 + C3D4 .inp file read by the reference's own reader/partitioner, integrated by femtech_b200 under the
    reference's symbol names (harness driver), vs the all-reference fixture mix4_p1."""
    exe = _need("dropin_ref_dump")
    from conftest import golden, rank_dict
    from oracle import pyoracle as po
    g = golden("mix4_p1")
    d = rank_dict(g, 0)
    etype = ["C3D8" if n == 8 else "C3D4" for n in np.diff(d["eptr"])]
    mesh.write_abaqus_inp_mixed(str(tmp_path / "mix.inp"), d["coordinates"].reshape(-1, 3), d["connectivity"], d["eptr"], d["pid"], etype)
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), d["materialID"], d["properties"])
    r = subprocess.run([exe, "mix.inp", "out", "150", repr(float(g["param_tMax"])), repr(float(g["param_dMax"]))],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    o = po.read_ref_dump(str(tmp_path / "out.rank0.bin"))
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert int(o["steps"][0]) == int(d["steps"][0]) and np.array_equal(o["eptr"], d["eptr"])
    assert rel(o["displacements"], d["displacements"]) < 1e-9 and rel(o["mass"], d["mass"]) < 1e-13
    assert o["F"].size == d["F"].size and rel(o["F"], d["F"]) < 1e-9 and rel(o["pk2"], d["pk2"]) < 1e-9
    assert rel(o["Eavg"], d["Eavg"]) < 1e-9


def test_resident_driver_one_explicit_dynamics_call(tmp_path):
    """integration/resident_driver.cpp: the reference's setup calls (its own reader, partitioner and allocator), the
    benchmark's boundary condition as a descriptor, then ONE ExplicitDynamics(tMax) through the reference's symbol --
    the whole loop on the GPU.  End state against the oracle at 1e-9; the energy file must hold one line per step
    (written from the records the device streams into the pinned step ring) with the oracle's running energies."""
    exe = _need("dropin_resident")
    from oracle import pyoracle as po
    X, conn, pid = mesh.cube_mesh(20)
    soft = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]
    mesh.write_abaqus_inp(str(tmp_path / "cube10.inp"), X, conn, pid)
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), [1], soft)
    tMax = 0.1  # the benchmark's run on a 20^3 mesh: ~290 steps, two slices of the host layer's loop
    r = subprocess.run([exe, "cube10.inp", repr(tMax), repr(0.007 * tMax / 0.1)], cwd=tmp_path, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-2000:])
    o = po.OracleModel(X, conn, pid, [1], soft)
    o.ShapeFunctions()
    o.AssembleLumpedMass()
    kind, rate = mesh.benchmark_bc(X, dMax=0.007 * tMax / 0.1, tMax=tMax)
    n, _, eh = po.run_explicit([o], [kind], rate, tMax, 10 ** 6)
    lines = open(tmp_path / "cube10.inp.resident.txt").read().split()
    nN, T = int(lines[0]), float(lines[1])
    u = np.array([float(x) for x in lines[3:]])
    assert nN == X.shape[0] and abs(T - o.Time) <= 1e-11 * o.Time
    assert np.abs(u - o.displacements).max() < 1e-9 * np.abs(o.displacements).max()
    en = np.array([[float(x) for x in l.split()] for l in open(tmp_path / [f for f in os.listdir(tmp_path) if f.startswith("energy_")][0])
                   if not l.startswith("#")])
    assert en.shape == (n, 5)  # one line per step
    assert np.all(np.diff(en[:, 0]) > 0) and abs(en[-1, 0] - o.Time) <= 1e-6 * o.Time
    eh = np.asarray(eh)
    assert np.allclose(en[:, 1:4], eh[:, :3], rtol=5e-6, atol=1e-30)


# ------------------------------------------------------------------------------------------------------------------
# Several ranks through the reference's own API: `ftmpirun -np P <driver>` -- one process per rank, the
# reference's reader and ParMETIS partitioner unchanged, femtech_host.cpp creating one context per rank, the shared-node
# sums moved by the reference's own MPI_Isend / MPI_Irecv loop (legacy GetForce / AssembleLumpedMass and the resident
# ExplicitDynamics with the host transport) or by the peer-memory windows (resident, p2p transport).  Checked against the
# fixtures the all-reference build wrote on the same number of ranks: maps bit-exact, state 1e-9.
def _bench10_inputs(tmp_path):
    from conftest import golden, rank_dict
    d = rank_dict(golden("bench10_p1"), 0)
    mesh.write_abaqus_inp(str(tmp_path / "bench10.inp"), d["coordinates"].reshape(-1, 3), d["connectivity"].reshape(-1, 8), d["pid"])
    mesh.write_materials_dat(str(tmp_path / "materials.dat"), d["materialID"], d["properties"])


def _rel(a, b):
    return float(np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("P", [2, 4, 8])
def test_multirank_legacy_calls_through_reference_symbols(tmp_path, P):
    """The harness driver (the drivers' host loop calling GetForce / CalculateAccelerations / StableTimeStep / CheckEnergy by
    their reference names every step) on P ranks: partition, node maps and send lists equal to the reference's, end state
    of every rank within 1e-9 of the reference's dump of that rank."""
    exe, mpirun = _need("dropin_ref_dump"), _need("ftmpirun")
    from conftest import golden, rank_dict
    from oracle import pyoracle as po
    g = golden("bench10_p%d" % P)
    _bench10_inputs(tmp_path)
    r = subprocess.run([mpirun, "-np", str(P), exe, "bench10.inp", "out", str(10 ** 9), repr(float(g["param_tMax"])),
                        repr(float(g["param_dMax"]))], cwd=tmp_path, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2000:])
    for rank in range(P):
        d = rank_dict(g, rank)
        o = po.read_ref_dump(str(tmp_path / ("out.rank%d.bin" % rank)))
        for k in ("connectivity", "global_eid", "globalNodeID", "sendProcessID", "sendNeighbourCountCum", "sendNodeIndex"):
            assert np.array_equal(o[k], d[k]), (rank, k)
        assert int(o["steps"][0]) == int(d["steps"][0])
        assert _rel(o["mass"], d["mass"]) < 1e-13
        assert _rel(o["displacements"], d["displacements"]) < 1e-9 and _rel(o["velocities"], d["velocities"]) < 1e-9
        assert _rel(o["fi"], d["fi"]) < 1e-6 and _rel(o["accelerations"], d["accelerations"]) < 1e-6
        assert abs(float(o["Time"][0]) - float(d["Time"][0])) <= 1e-11 * float(d["Time"][0])


def _resident_multirank(tmp_path, P, transport):
    exe, mpirun = _need("dropin_resident"), _need("ftmpirun")
    from conftest import golden, rank_dict
    g = golden("bench10_p%d" % P)
    _bench10_inputs(tmp_path)
    env = dict(os.environ, FTB200_MPI_TRANSPORT=transport, CUDA_DEVICE_MAX_CONNECTIONS="32")
    r = subprocess.run([mpirun, "-np", str(P), exe, "bench10.inp", repr(float(g["param_tMax"])), repr(float(g["param_dMax"]))],
                       cwd=tmp_path, capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2000:])
    for rank in range(P):
        d = rank_dict(g, rank)
        vals = open(tmp_path / ("bench10.inp.resident.rank%d.txt" % rank)).read().split()
        nN = int(vals[0])
        assert nN == d["globalNodeID"].size
        u = np.array(vals[3:3 + 3 * nN], dtype=float)
        gid = np.array(vals[3 + 3 * nN:3 + 4 * nN], dtype=np.int64)
        v = np.array(vals[3 + 4 * nN:3 + 7 * nN], dtype=float)
        assert np.array_equal(gid, d["globalNodeID"])  # the reference's ParMETIS split and numbering, bit for bit
        assert abs(float(vals[1]) - float(d["Time"][0])) <= 1e-11 * float(d["Time"][0])
        assert _rel(u, d["displacements"]) < 1e-9 and _rel(v, d["velocities"]) < 1e-9
    en = np.array([[float(x) for x in l.split()] for l in open(tmp_path / [f for f in os.listdir(tmp_path) if f.startswith("energy_")][0])
                   if not l.startswith("#")])
    assert en.shape[0] == int(rank_dict(g, 0)["steps"][0])  # one line per step, written by rank 0 from the summed records
    ef = g["energy_file"][-1]
    assert np.allclose(en[-1, 1:4], ef[1:4], rtol=5e-6, atol=1e-30)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_multirank_resident_explicit_dynamics_host_transport(tmp_path, P):
    """integration/resident_driver.cpp on P ranks: ONE ExplicitDynamics call per rank, one partition per context, the
    per-step exchange through the reference's MPI loop (device packs and adds)."""
    _resident_multirank(tmp_path, P, "host")


def test_multirank_resident_explicit_dynamics_peer_memory(tmp_path):
    """Same with the peer-memory windows (CUDA IPC handles exchanged with MPI_Allgather, the loop as a CUDA graph, no MPI per
    step).  Needs one GPU per rank: processes that share a device time-slice, and a kernel waiting for a peer's flag would
    burn the peer's time slice."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("peer-memory transport across processes needs one GPU per rank")
    _resident_multirank(tmp_path, 2, "p2p")
