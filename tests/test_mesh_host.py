"""Host-side logic: synthetic mesh generator and BC descriptor (no GPU)."""
import numpy as np

from femtech_b200 import mesh
from conftest import golden, rank_dict


def test_cube_mesh_counts_and_orientation():
    X, conn, pid = mesh.cube_mesh(5, nparts_z=2)
    assert X.shape == (216, 3) and conn.shape == (125, 8) and set(pid) == {0, 1}
    # positive Jacobian: (n1-n0) x (n3-n0) . (n4-n0) > 0
    a, b, c = X[conn[:, 1]] - X[conn[:, 0]], X[conn[:, 3]] - X[conn[:, 0]], X[conn[:, 4]] - X[conn[:, 0]]
    assert np.all(np.einsum("ij,ij->i", np.cross(a, b), c) > 0)
    assert X.max() == mesh.CUBE_L and X.min() == 0.0


def test_benchmark_bc_matches_reference_boundary_flags():
    """bc descriptor reproduces the boundary[] the reference driver set on the shipped mesh."""
    d = rank_dict(golden("bench10_p1"), 0)
    kind, rate = mesh.benchmark_bc(d["coordinates"])
    assert np.array_equal((kind > 0).astype(np.int32), d["boundary0"])
    assert rate[2] == 0.007 / 0.1


def test_inp_roundtrip_is_exact(tmp_path):
    """The reference reader ingested our .inp bit-exactly (coordinates survived %.17g)."""
    d = rank_dict(golden("cube4j_m1"), 0)
    X, conn, pid = mesh.cube_mesh(4, jitter=0.1)
    # the reference renumbers nodes by sorted global id == our ids, so arrays must be identical
    assert np.array_equal(d["coordinates"], X.reshape(-1))
    assert np.array_equal(d["connectivity"], conn.reshape(-1))
