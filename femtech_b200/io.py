"""On-disk formats either side of the hot path, host side.

Readers follow the reference's semantics so that the same file yields the same arrays the reference holds after
ReadInputFile + PartitionMesh on one rank:
  read_abaqus_inp   src/io/input/ReadAbaqus.cpp:26-277     *NODE / *ELEMENT, TYPE=..., ELSET=...; part id = order of
                                                           first appearance of the ELSET name (:171-178)
  read_lsdyna_k     src/io/input/ReadLsDyna.cpp:48-321     *ELEMENT_SOLID (eid pid n1..n8; C3D4 when 4 unique nodes,
                                                           :224-235), *NODE; part id = pid - 1 (:223)
  ReadInputFile     src/io/input/ReadInputFile.cpp         dispatch on the extension
  ReadMaterials     src/io/input/ReadMaterials.cpp:8-138   `partID materialID rho [mu lambda [k1 k2 [g1 t1 g2 t2]]]`
  localize          src/io/PartitionMesh.cpp:485-536       local node numbering (ascending global id), 1 rank
Writers produce what ParaView reads from the reference (src/io/output/WriteVTU.cpp:3-263, WritePVD.cpp): the same
piece layout and array names (Displacements, Accelerations, Boundary | PartID, AvgStrain, ProcID, extra int cell
arrays), either in the reference's ASCII form or -- the default, since a 100^3 step takes 0.3 ms on the device and an
ASCII dump of it seconds -- as raw appended binary.
"""
import os
import re

import numpy as np

_NUM = re.compile(r"^[\s,]*[-+]?(\d|\.\d)")
NODES_OF = {"C3D8": 8, "C3D4": 4, "T3D2": 2}
VTK_TYPE = {"C3D8": 12, "C3D4": 10, "T3D2": 3}  # WriteVTU.cpp:76-87


def _is_data(line):
    return bool(_NUM.match(line))


def read_abaqus_inp(path):
    """-> dict(node_ids, node_xyz, elem_ids, elem_type[list], conn[list of arrays, 0-based GLOBAL node ids], pid, elsets)"""
    node_ids, node_xyz, elem_ids, etype, conn, set_of = [], [], [], [], [], []
    elsets = []
    mode, cur_type, cur_set = None, None, None
    nodes_done = False
    with open(path, "r", errors="replace") as f:
        for line in f:
            s = line.strip()
            if s.startswith("*"):
                if s.startswith("**"):
                    continue
                head = [t.strip() for t in s[1:].split(",")]
                key = head[0].upper()
                mode = None
                if key == "NODE" and not nodes_done:  # only the first node section is used (:41)
                    mode = "node"
                elif key == "ELEMENT":
                    opts = {}
                    for t in head[1:]:
                        if "=" in t:
                            k, v = t.split("=", 1)
                            opts[k.strip().upper()] = v.strip()
                    if "TYPE" in opts:  # IsElementSection requires TYPE (:389-394)
                        mode, cur_type, cur_set = "elem", opts["TYPE"].upper(), opts.get("ELSET", "")
                        if cur_set not in elsets:
                            elsets.append(cur_set)
                continue
            if mode is None or not s:
                continue
            if not _is_data(s):
                mode = None
                continue
            vals = [v for v in re.split(r"[,\s]+", s) if v]
            if mode == "node":
                node_ids.append(int(float(vals[0])))
                xyz = [float(v) for v in vals[1:4]]
                node_xyz.append(xyz + [0.0] * (3 - len(xyz)))
                nodes_done = True
            else:
                ints = [int(v) for v in vals]
                elem_ids.append(ints[0])
                conn.append(np.array(ints[1:], dtype=np.int64) - 1)
                etype.append(cur_type)
                set_of.append(elsets.index(cur_set))
    if not elem_ids:
        raise ValueError("%s: no element found" % path)
    if not node_ids:
        raise ValueError("%s: node section not found" % path)
    return dict(node_ids=np.array(node_ids, dtype=np.int64), node_xyz=np.array(node_xyz, dtype=np.float64),
                elem_ids=np.array(elem_ids, dtype=np.int32), elem_type=etype, conn=conn,
                pid=np.array(set_of, dtype=np.int32), elsets=elsets)


def read_lsdyna_k(path):
    node_ids, node_xyz, elem_ids, etype, conn, pids = [], [], [], [], [], []
    mode = None
    with open(path, "r", errors="replace") as f:
        for line in f:
            s = line.rstrip("\n")
            if s.startswith("$"):
                continue
            if s.startswith("*"):
                key = s.strip().upper()
                mode = "elem" if key.startswith("*ELEMENT_SOLID") or key.startswith("*ELEMENT_BEAM") or \
                    key.startswith("*ELEMENT_SHELL") else ("node" if key == "*NODE" else None)
                continue
            if mode is None or not s.strip():
                continue
            vals = s.split()
            if mode == "node":
                node_ids.append(int(vals[0]))
                xyz = [float(v) for v in vals[1:4]]
                node_xyz.append(xyz + [0.0] * (3 - len(xyz)))
            else:
                ints = [int(v) for v in vals]
                nodes = np.array(ints[2:], dtype=np.int64)
                nuniq = len(np.unique(nodes)) if len(nodes) == 8 else len(nodes)  # ReadLsDyna.cpp:224
                t = "C3D8" if nuniq > 4 else ("C3D4" if nuniq == 4 else "T3D2")
                elem_ids.append(ints[0])
                pids.append(ints[1] - 1)
                etype.append(t)
                conn.append(nodes - 1)  # all columns are kept, like the reference (:238-241)
    if not elem_ids:
        raise ValueError("%s: no element found" % path)
    return dict(node_ids=np.array(node_ids, dtype=np.int64), node_xyz=np.array(node_xyz, dtype=np.float64),
                elem_ids=np.array(elem_ids, dtype=np.int32), elem_type=etype, conn=conn,
                pid=np.array(pids, dtype=np.int32), elsets=None)


def ReadInputFile(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".k":
        return read_lsdyna_k(path)
    if ext == ".inp":
        return read_abaqus_inp(path)
    raise ValueError("unknown mesh file extension %r (the reference accepts .k and .inp)" % ext)


def localize(mesh):
    """One-rank PartitionMesh: nodes used by the elements, local id = rank of the global id (PartitionMesh.cpp:485-536).
    -> coordinates [3*nNodes], connectivity (flat, local ids), eptr, pid, globalNodeID (1-based), global_eid."""
    flat = np.concatenate(mesh["conn"])
    gids = np.unique(flat)
    local = np.searchsorted(gids, flat).astype(np.int32)
    eptr = np.concatenate([[0], np.cumsum([len(c) for c in mesh["conn"]])]).astype(np.int32)
    order = np.argsort(mesh["node_ids"], kind="stable")
    pos = np.searchsorted(mesh["node_ids"][order], gids + 1)
    if np.any(pos >= len(order)) or np.any(mesh["node_ids"][order][pos] != gids + 1):
        raise ValueError("element references a node that the node section does not define")
    X = mesh["node_xyz"][order][pos]
    return dict(coordinates=np.ascontiguousarray(X.reshape(-1)), connectivity=local, eptr=eptr, pid=mesh["pid"].copy(),
                globalNodeID=(gids + 1).astype(np.int32), global_eid=mesh["elem_ids"].copy(),
                ElementType=list(mesh["elem_type"]))


def ReadMaterials(path, nPID):
    """-> materialID [nPID], properties [9*nPID] (rho mu lambda k1 k2 g1 t1 g2 t2; unused entries 0)."""
    nvals = {0: 1, 1: 3, 2: 3, 3: 3, 4: 5, 5: 9}
    mat = np.zeros(nPID, dtype=np.int32)
    props = np.zeros(9 * nPID)
    seen = 0
    with open(path) as f:
        for line in f:
            v = line.split()
            if len(v) < 2:
                continue
            p, m = int(v[0]), int(v[1])
            if m not in nvals:
                raise ValueError("materials.dat: unknown material %d for part %d" % (m, p))
            if p >= nPID:
                continue
            mat[p] = m
            vals = [float(x) for x in v[2:2 + nvals[m]]]
            if len(vals) < nvals[m]:
                raise ValueError("materials.dat: part %d needs %d values" % (p, nvals[m]))
            props[9 * p:9 * p + len(vals)] = vals
            seen += 1
    if seen < nPID:
        raise ValueError("materials.dat defines %d of %d parts" % (seen, nPID))
    return mat, props


# ------------------------------------------------------------------------------------------------ writers
def _ascii_rows(a, fmt, per_row):
    a = np.asarray(a).reshape(-1, per_row)
    return "".join("\t\t\t\t\t" + " ".join(fmt % x for x in row) + " \n" for row in a)


def WriteVTU(path, coordinates, displacements, connectivity, eptr, ElementType, pid, accelerations=None, boundary=None,
             Eavg=None, rank=0, int_cell_data=None, binary=True):
    """One piece (WriteVTU.cpp:31-206).  binary=True: raw appended data (UInt64 headers); False: the reference's ASCII."""
    X = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    U = np.asarray(displacements, dtype=np.float64).reshape(-1, 3)
    U = np.where(np.abs(U) < 1e-20, 0.0, U)  # WriteVTU.cpp:41-43
    nN, nE = X.shape[0], len(eptr) - 1
    arrays = []  # (section, name, dtype, ncomp, data)
    arrays.append(("Points", None, "Float64", 3, X + U))
    arrays.append(("Cells", "connectivity", "Int32", 1, np.asarray(connectivity, dtype=np.int32)))
    arrays.append(("Cells", "offsets", "Int32", 1, np.asarray(eptr[1:], dtype=np.int32)))
    arrays.append(("Cells", "types", "Int32", 1, np.array([VTK_TYPE[t] for t in ElementType], dtype=np.int32)))
    arrays.append(("PointData", "Displacements", "Float64", 3, U))
    if accelerations is not None:
        arrays.append(("PointData", "Accelerations", "Float64", 3, np.asarray(accelerations, dtype=np.float64).reshape(-1, 3)))
    if boundary is not None:
        arrays.append(("PointData", "Boundary", "Int32", 3, np.asarray(boundary, dtype=np.int32).reshape(-1, 3)))
    arrays.append(("CellData", "PartID", "Int32", 1, np.asarray(pid, dtype=np.int32)))
    if Eavg is not None:
        arrays.append(("CellData", "AvgStrain", "Float64", 9, np.asarray(Eavg, dtype=np.float64).reshape(-1, 9)))
    arrays.append(("CellData", "ProcID", "Int32", 1, np.full(nE, rank, dtype=np.int32)))
    for name, data in (int_cell_data or {}).items():
        arrays.append(("CellData", name, "Int32", 1, np.asarray(data, dtype=np.int32)))
    out = ['<?xml version="1.0"?>\n',
           '<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian"%s>\n'
           % (' header_type="UInt64"' if binary else ""),
           "\t<UnstructuredGrid>\n", '\t\t<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (nN, nE)]
    attrs = {"PointData": ' Vectors="Displacements Accelerations"', "CellData": ' Scalars="PartID" Tensors="AvgStrain"'}
    blobs, offset, cur = [], 0, None
    for sec, name, typ, nc, data in arrays:
        if sec != cur:
            if cur is not None:
                out.append("\t\t\t</%s>\n" % cur)
            out.append("\t\t\t<%s%s>\n" % (sec, attrs.get(sec, "")))
            cur = sec
        tag = '\t\t\t\t<DataArray type="%s"%s NumberOfComponents="%d"' % (typ, ' Name="%s"' % name if name else "", nc)
        if binary:
            raw = np.ascontiguousarray(data).tobytes()
            out.append(tag + ' format="appended" offset="%d"/>\n' % offset)
            blobs.append(np.uint64(len(raw)).tobytes() + raw)
            offset += 8 + len(raw)
        else:
            fmt = "%10.8e" if typ == "Float64" else "%d"
            per = nc if nc > 1 else 1
            out.append(tag + ' format="ascii">\n' + _ascii_rows(data, fmt, per) + "\t\t\t\t</DataArray>\n")
    out.append("\t\t\t</%s>\n\t\t</Piece>\n\t</UnstructuredGrid>\n" % cur)
    with open(path, "wb") as f:
        f.write("".join(out).encode())
        if binary:
            f.write(b'\t<AppendedData encoding="raw">\n_')
            for b in blobs:
                f.write(b)
            f.write(b"\n\t</AppendedData>\n")
        f.write(b"</VTKFile>\n")


def WritePVTU(path, piece_files, extra_cell_arrays=()):
    """WriteVTU.cpp:208-260: the parallel index of the pieces."""
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        f.write('\t<PUnstructuredGrid GhostLevel="0">\n\t\t<PPoints>\n')
        f.write('\t\t\t<PDataArray type="Float64" NumberOfComponents="3"/>\n\t\t</PPoints>\n')
        f.write('\t\t<PPointData Vectors="Displacements Accelerations">\n')
        f.write('\t\t\t<PDataArray type="Float64" Name="Displacements" NumberOfComponents="3"/>\n')
        f.write('\t\t\t<PDataArray type="Float64" Name="Accelerations" NumberOfComponents="3"/>\n')
        f.write('\t\t\t<PDataArray type="Int32" Name="Boundary" NumberOfComponents="3"/>\n\t\t</PPointData>\n')
        f.write('\t\t<PCellData Scalars="PartID" Tensors="AvgStrain">\n\t\t\t<PDataArray type="Int32" Name="PartID"/>\n')
        f.write('\t\t\t<PDataArray type="Float64" Name="AvgStrain" NumberOfComponents="9"/>\n')
        f.write('\t\t\t<PDataArray type="Int32" Name="ProcID"/>\n')
        for n in extra_cell_arrays:
            f.write('\t\t\t<PDataArray type="Int32" Name="%s"/>\n' % n)
        f.write("\t\t</PCellData>\n")
        for p in piece_files:
            f.write('\t\t<Piece Source="%s"/>\n' % p)
        f.write("\t</PUnstructuredGrid>\n</VTKFile>\n")


def WritePVD(path, times, files):
    """WritePVD.cpp:25-34: the time collection."""
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n\t<Collection>\n')
        for t, fn in zip(times, files):
            f.write('\t\t<DataSet timestep="%f" file="%s" />\n' % (t, fn))
        f.write("\t</Collection>\n</VTKFile>\n")


def read_vtu_arrays(path):
    """Minimal reader of the files WriteVTU produces (tests, post-processing): name -> ndarray."""
    raw = open(path, "rb").read()
    head_end = raw.find(b"<AppendedData")
    text = raw[:head_end if head_end >= 0 else len(raw)].decode()
    out = {}
    np_t = {"Float64": np.float64, "Int32": np.int32}
    if head_end >= 0:
        base = raw.index(b"_", head_end) + 1
        for m in re.finditer(r'<DataArray type="(\w+)"(?: Name="([\w.-]+)")? NumberOfComponents="(\d+)" format="appended" offset="(\d+)"/>', text):
            typ, name, nc, off = m.group(1), m.group(2) or "Points", int(m.group(3)), int(m.group(4))
            n = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=base + off)[0])
            a = np.frombuffer(raw, dtype=np_t[typ], count=n // np.dtype(np_t[typ]).itemsize, offset=base + off + 8)
            out[name] = a.reshape(-1, nc) if nc > 1 else a.copy()
    else:
        for m in re.finditer(r'<DataArray type="(\w+)"(?: Name="([\w.-]+)")? NumberOfComponents="(\d+)" format="ascii">\n(.*?)</DataArray>',
                             text, flags=re.S):
            typ, name, nc = m.group(1), m.group(2) or "Points", int(m.group(3))
            a = np.array(m.group(4).split(), dtype=np_t[typ])
            out[name] = a.reshape(-1, nc) if nc > 1 else a
    return out
