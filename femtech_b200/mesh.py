"""Synthetic hex8 meshes, the benchmark boundary condition and an Abaqus
``.inp`` writer (so the reference reader ingests the identical mesh).

Follows SURVEY.md section 8(d): cube [0,L]^3, L = 0.005 m, node (i,j,k) at
(i*h, j*h, k*h), global node id = i + (n+1)*j + (n+1)^2*k, element (i,j,k) id =
i + n*j + n^2*k with C3D8 ordering [n000,n100,n110,n010,n001,n101,n111,n011]
(reference: src/fem/ShapeFunctions/ShapeFunction_C3D8.cpp:22-29).
"""
import numpy as np

CUBE_L = 0.005  # examples/Benchmarking-Parallel/*.inp side length


def cube_mesh(n, L=CUBE_L, jitter=0.0, seed=1234, nparts_z=1):
    """Structured n^3 hex8 cube.

    Returns (coordinates[N,3] f64, connectivity[E,8] i32, pid[E] i32).
    ``jitter`` > 0 perturbs interior nodes by U(-jitter*h, jitter*h) per
    coordinate with PCG64(seed) (exercises the non-parallelogram face branch of
    src/math/Geometry.cpp:46-63).  ``nparts_z`` splits the cube into that many
    z-slabs with part ids 0..nparts_z-1 (multi-material meshes).
    """
    h = L / n
    g = np.arange(n + 1, dtype=np.float64) * h
    # node id = i + (n+1) j + (n+1)^2 k  -> k slowest
    kk, jj, ii = np.meshgrid(np.arange(n + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    X = np.stack([g[ii], g[jj], g[kk]], axis=-1).reshape(-1, 3)
    if jitter > 0.0:
        rng = np.random.Generator(np.random.PCG64(seed))
        interior = ((ii > 0) & (ii < n) & (jj > 0) & (jj < n) & (kk > 0) & (kk < n)).reshape(-1)
        d = rng.uniform(-jitter * h, jitter * h, size=X.shape)
        X[interior] += d[interior]
    ek, ej, ei = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ei, ej, ek = ei.reshape(-1), ej.reshape(-1), ek.reshape(-1)
    n1 = n + 1

    def nid(i, j, k):
        return (i + n1 * j + n1 * n1 * k).astype(np.int32)

    conn = np.stack([nid(ei, ej, ek), nid(ei + 1, ej, ek), nid(ei + 1, ej + 1, ek), nid(ei, ej + 1, ek),
                     nid(ei, ej, ek + 1), nid(ei + 1, ej, ek + 1), nid(ei + 1, ej + 1, ek + 1),
                     nid(ei, ej + 1, ek + 1)], axis=1).astype(np.int32)
    pid = np.minimum((ek * nparts_z) // n, nparts_z - 1).astype(np.int32)
    return np.ascontiguousarray(X), np.ascontiguousarray(conn), pid


def box_mesh(Nx, Ny, Nz, h):
    """Structured Nx x Ny x Nz hex8 box with spacing h, numbered like cube_mesh (node id = i + (Nx+1) j + (Nx+1)(Ny+1) k,
    element id = i + Nx j + Nx Ny k): the global mesh of femtech_b200.dist.brick_partition's per-rank bricks."""
    kk, jj, ii = np.meshgrid(np.arange(Nz + 1), np.arange(Ny + 1), np.arange(Nx + 1), indexing="ij")
    X = np.stack([ii.reshape(-1) * h, jj.reshape(-1) * h, kk.reshape(-1) * h], axis=-1).astype(np.float64)
    ek, ej, ei = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij")
    ei, ej, ek = ei.reshape(-1), ej.reshape(-1), ek.reshape(-1)
    nx1, nxy = Nx + 1, (Nx + 1) * (Ny + 1)

    def nid(i, j, k):
        return (i + nx1 * j + nxy * k).astype(np.int32)

    conn = np.stack([nid(ei, ej, ek), nid(ei + 1, ej, ek), nid(ei + 1, ej + 1, ek), nid(ei, ej + 1, ek),
                     nid(ei, ej, ek + 1), nid(ei + 1, ej, ek + 1), nid(ei + 1, ej + 1, ek + 1),
                     nid(ei, ej + 1, ek + 1)], axis=1).astype(np.int32)
    return np.ascontiguousarray(X), np.ascontiguousarray(conn), np.zeros(conn.shape[0], np.int32)


KUHN_TETS = ((0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6))


def split_hex_to_tets(conn, pid, which):
    """Mixed C3D8 / C3D4 mesh: hexahedra with which[e] true are replaced by six tetrahedra around the 0-6 diagonal
    (conforming between neighbours of a structured mesh).  Returns (connectivity packed 8 or 4 per element [flat],
    eptr[E'+1], pid[E'], etype list) in element order: every hexahedron stays in place, its tets take its slot."""
    flat, eptr, pids, etype = [], [0], [], []
    for e in range(conn.shape[0]):
        if which[e]:
            for t in KUHN_TETS:
                flat.extend(int(conn[e, k]) for k in t)
                eptr.append(eptr[-1] + 4)
                pids.append(int(pid[e]))
                etype.append("C3D4")
        else:
            flat.extend(int(v) for v in conn[e])
            eptr.append(eptr[-1] + 8)
            pids.append(int(pid[e]))
            etype.append("C3D8")
    return np.array(flat, dtype=np.int32), np.array(eptr, dtype=np.int32), np.array(pids, dtype=np.int32), etype


def write_abaqus_inp_mixed(path, coordinates, connectivity, eptr, pid, etype):
    """write_abaqus_inp for mixed element types: one *ELEMENT block per run of equal (part, type), so the file order is
    the element order and part ids follow the first appearance of the ELSET names (ReadAbaqus.cpp:171-178)."""
    X = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    order = []
    for p in pid:
        if p not in order:
            order.append(int(p))
    if order != sorted(order):
        raise ValueError("part ids must first appear in ascending order (ReadAbaqus.cpp:173-179)")
    with open(path, "w") as f:
        f.write("*Heading\n** femtech_b200 synthetic mixed mesh\n*Node\n")
        for i, (x, y, z) in enumerate(X):
            f.write("%d, %.17g, %.17g, %.17g\n" % (i + 1, x, y, z))
        prev = None
        for e in range(len(pid)):
            key = (int(pid[e]), etype[e])
            if key != prev:
                f.write("*ELEMENT,TYPE=%s,ELSET=PART_%d\n" % (etype[e], key[0] + 1))
                prev = key
            nodes = connectivity[eptr[e]:eptr[e + 1]]
            f.write("%d, %s\n" % (e + 1, ", ".join(str(int(v) + 1) for v in nodes)))


def benchmark_bc(coordinates, L=CUBE_L, dMax=0.007, tMax=0.1, tol=1e-5):
    """The benchmark driver's boundary condition as a descriptor.

    examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp:184-244: u_x = 0 on
    x = 0, u_y = 0 on y = 0, u_z = 0 on z = 0, u_y = Time*dMax/tMax and
    v_y = dMax/tMax on y = L.  Returns (bc_kind[3N] int32, bc_rate[3] f64):
    kind 0 = free, kind k > 0 -> u = Time*bc_rate[k], v = bc_rate[k], a = 0.
    """
    X = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    kind = np.zeros(X.shape, dtype=np.int32)
    kind[np.abs(X - 0.0) < tol] = 1
    kind[np.abs(X[:, 1] - L) < tol, 1] = 2
    rate = np.array([0.0, 0.0, dMax / tMax], dtype=np.float64)
    return kind.reshape(-1), rate


def write_abaqus_inp(path, coordinates, connectivity, pid=None):
    """Write a mesh the reference's reader accepts (src/io/input/ReadAbaqus.cpp:
    26-277): one *Node block, one *Element block per part in order of first
    appearance, 1-based ids, coordinates with 17 significant digits (the reader
    is strtod based, ReadInputFile.cpp:104, so they round-trip exactly)."""
    X = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    conn = np.asarray(connectivity).reshape(-1, 8)
    if pid is None:
        pid = np.zeros(conn.shape[0], dtype=np.int32)
    with open(path, "w") as f:
        f.write("*Heading\n** femtech_b200 synthetic hex8 mesh\n*Node\n")
        for i, (x, y, z) in enumerate(X):
            f.write("%d, %.17g, %.17g, %.17g\n" % (i + 1, x, y, z))
        order = []
        for p in pid:
            if p not in order:
                order.append(int(p))
        if order != sorted(order):
            raise ValueError("part ids must first appear in ascending order (ReadAbaqus.cpp:173-179)")
        eid = 0
        # elements must stay in file order == element order, so emit runs of equal pid
        start = 0
        E = conn.shape[0]
        while start < E:
            end = start
            while end < E and pid[end] == pid[start]:
                end += 1
            f.write("*ELEMENT,TYPE=C3D8,ELSET=PART_%d\n" % (int(pid[start]) + 1))
            for e in range(start, end):
                eid += 1
                f.write("%d, %s\n" % (eid, ", ".join(str(int(v) + 1) for v in conn[e])))
            start = end
        f.write("*End\n")


def write_materials_dat(path, materialID, properties):
    """materials.dat in the layout of src/io/input/ReadMaterials.cpp:43-122."""
    nvals = {0: 1, 1: 3, 2: 3, 3: 3, 4: 5, 5: 9}
    props = np.asarray(properties, dtype=np.float64).reshape(-1, 9)
    with open(path, "w") as f:
        for p, m in enumerate(materialID):
            vals = " ".join("%.17g" % v for v in props[p, :nvals[int(m)]])
            f.write("%d %d %s\n" % (p, int(m), vals))
