"""On-disk formats: the readers give the arrays the reference holds after ReadInputFile +
PartitionMesh on one rank (checked against the fixtures dumped from the reference for its own shipped mesh files, when
those files are present in this container), and the writers round-trip."""
import os

import numpy as np
import pytest

from conftest import golden, rank_dict
from femtech_b200 import io as fio
from femtech_b200 import mesh

REF = os.environ.get("FEMTECH_REFERENCE", "/root/reference")


def test_abaqus_writer_reader_round_trip(tmp_path):
    X, conn, pid = mesh.cube_mesh(5, jitter=0.1, nparts_z=3)
    p = str(tmp_path / "c.inp")
    mesh.write_abaqus_inp(p, X, conn, pid)
    m = fio.localize(fio.ReadInputFile(p))
    assert np.allclose(m["coordinates"].reshape(-1, 3), X, rtol=0, atol=1e-16 + 1e-15 * np.abs(X).max())
    assert np.array_equal(m["connectivity"].reshape(-1, 8), conn)
    assert np.array_equal(m["pid"], pid)
    assert np.array_equal(m["eptr"], 8 * np.arange(conn.shape[0] + 1))
    assert set(m["ElementType"]) == {"C3D8"}


def test_lsdyna_reader_and_degenerate_tet(tmp_path):
    p = tmp_path / "m.k"
    p.write_text("*KEYWORD\n*ELEMENT_SOLID\n$# eid pid n1..n8\n 7 2 1 2 4 3 5 6 8 7\n 9 1 1 2 3 5 5 5 5 5\n*NODE\n$# nid x y z\n"
                 + "".join("%d %g %g %g 0 0\n" % (i + 1, i & 1, (i >> 1) & 1, (i >> 2) & 1) for i in range(8)) + "*END\n")
    r = fio.ReadInputFile(str(p))
    assert r["elem_type"] == ["C3D8", "C3D4"] and list(r["pid"]) == [1, 0] and list(r["elem_ids"]) == [7, 9]
    assert len(r["conn"][1]) == 8  # the reference keeps all eight columns of a degenerate solid (ReadLsDyna.cpp:238-241)


@pytest.mark.parametrize("fixture,relpath", [("bench10_p1", "examples/Benchmarking-Parallel/10elements.inp"),
                                             ("ex9_1elt", "examples/ex9/1-elt-cube.k")])
def test_readers_match_the_reference_on_its_shipped_meshes(fixture, relpath):
    path = os.path.join(REF, relpath)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    d = rank_dict(golden(fixture), 0)
    m = fio.localize(fio.ReadInputFile(path))
    assert np.array_equal(m["coordinates"], d["coordinates"])
    assert np.array_equal(m["connectivity"], d["connectivity"])
    assert np.array_equal(m["pid"], d["pid"])
    if "globalNodeID" in d:
        assert np.array_equal(m["globalNodeID"], d["globalNodeID"]) and np.array_equal(m["global_eid"], d["global_eid"])


def test_materials_round_trip_and_errors(tmp_path):
    p = str(tmp_path / "materials.dat")
    brain = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]
    props = [1500.0, 0, 0, 0, 0, 0, 0, 0, 0] + [1040.0, 100.0, 200.0, 0, 0, 0, 0, 0, 0] + brain
    mesh.write_materials_dat(p, [0, 1, 5], props)
    mat, pr = fio.ReadMaterials(p, 3)
    assert list(mat) == [0, 1, 5] and np.array_equal(pr, np.array(props))
    with pytest.raises(ValueError):
        fio.ReadMaterials(p, 4)
    (tmp_path / "bad.dat").write_text("0 7 1000\n")
    with pytest.raises(ValueError):  # StressUpdate.cpp:24 / ReadMaterials.cpp: unknown material
        fio.ReadMaterials(str(tmp_path / "bad.dat"), 1)


@pytest.mark.parametrize("binary", [True, False])
def test_vtu_round_trip(tmp_path, binary):
    X, conn, pid = mesh.cube_mesh(3, nparts_z=3)
    rng = np.random.default_rng(2)
    U = 1e-4 * rng.standard_normal(X.shape)
    U[0] = 1e-30  # flushed to zero like WriteVTU.cpp:41-43
    A = rng.standard_normal(X.shape)
    B = (rng.random(X.shape) < 0.2).astype(np.int32)
    E = rng.standard_normal((conn.shape[0], 9))
    flag = (rng.random(conn.shape[0]) < 0.5).astype(np.int32)
    p = str(tmp_path / "out.vtu")
    fio.WriteVTU(p, X.reshape(-1), U.reshape(-1), conn.reshape(-1), 8 * np.arange(conn.shape[0] + 1), ["C3D8"] * conn.shape[0],
                 pid, accelerations=A.reshape(-1), boundary=B.reshape(-1), Eavg=E.reshape(-1), rank=3,
                 int_cell_data={"CSDM-15": flag}, binary=binary)
    a = fio.read_vtu_arrays(p)
    tol = 0 if binary else 1e-8
    Uz = U.copy()
    Uz[0] = 0.0
    assert np.allclose(a["Points"], X + Uz, rtol=tol, atol=tol * 1e-3)
    assert np.array_equal(a["connectivity"], conn.reshape(-1)) and np.array_equal(a["offsets"], 8 * np.arange(1, conn.shape[0] + 1))
    assert set(a["types"]) == {12} and np.array_equal(a["PartID"], pid) and set(a["ProcID"]) == {3}
    assert np.allclose(a["Displacements"], Uz, rtol=tol, atol=1e-12) and np.allclose(a["AvgStrain"], E, rtol=1e-8)
    assert np.array_equal(a["Boundary"], B) and np.array_equal(a["CSDM-15"], flag)
    fio.WritePVTU(str(tmp_path / "out.pvtu"), ["out.vtu"], ["CSDM-15"])
    fio.WritePVD(str(tmp_path / "out.pvd"), [0.0, 1e-3], ["out.0000.pvtu", "out.0001.pvtu"])
    assert 'timestep="0.001000"' in open(tmp_path / "out.pvd").read()
