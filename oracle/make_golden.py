#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  TEST INFRASTRUCTURE.

Runs oracle/_ref/ref_dump_exact (the unmodified reference sources compiled by
oracle/ref/build_ref.sh, driven by oracle/ref/ref_dump.cpp) on the shipped
meshes and on small synthetic ones, single- and multi-rank (through the ftmpi
shim), and stores compact fixtures.  Must be run in the build container
(needs /root/reference); the fixtures travel, the reference does not.

    python oracle/make_golden.py            # regenerate everything
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femtech_b200 import mesh  # noqa: E402
from oracle import pyoracle  # noqa: E402

REF = os.environ.get("FEMTECH_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]  # examples/ex5/materials.dat:2
SOFT = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]  # examples/Benchmarking-Parallel/materials.dat:1


def write_rigid_tables(path, tables):
    """6 lines `n t0 v0 t1 v1 ...`: angular x,y,z then linear x,y,z acceleration traces (s, SI), ref_dump REF_RIGID."""
    with open(path, "w") as f:
        for t, v in tables:
            f.write("%d %s\n" % (len(t), " ".join("%.17g %.17g" % (a, b) for a, b in zip(t, v))))


def run_ref(meshfile, materialID, properties, nranks, maxSteps, tMax, dMax, cubeL=mesh.CUBE_L, injury_exclude=None,
            rigid_tables=None):
    work = tempfile.mkdtemp(prefix="ftgold_")
    env = dict(os.environ)
    if injury_exclude is not None:
        env["REF_INJURY_EXCLUDE"] = ",".join(str(p) for p in injury_exclude)
    if rigid_tables is not None:
        write_rigid_tables(os.path.join(work, "rigid.txt"), rigid_tables)
        env["REF_RIGID"] = "rigid.txt"
    try:
        mesh.write_materials_dat(os.path.join(work, "materials.dat"), materialID, properties)
        cmd = [os.path.join(BIN, "ref_dump_exact"), meshfile, os.path.join(work, "out"), str(maxSteps),
               repr(tMax), repr(dMax), repr(cubeL)]
        if nranks > 1:
            cmd = [os.path.join(BIN, "ftmpirun"), "-np", str(nranks)] + cmd
        out = subprocess.run(cmd, cwd=work, check=True, capture_output=True, text=True, env=env).stdout
        dumps = [pyoracle.read_ref_dump(os.path.join(work, "out.rank%d.bin" % r)) for r in range(nranks)]
        energy = None
        for fn in os.listdir(work):
            if fn.startswith("energy_"):
                rows = [l.split() for l in open(os.path.join(work, fn)) if not l.startswith("#")]
                energy = np.array(rows, dtype=np.float64) if rows else np.zeros((0, 5))
        return dumps, energy, out
    finally:
        shutil.rmtree(work, ignore_errors=True)


MESH_KEYS = ["coordinates", "connectivity", "pid", "materialID", "properties"]
MAP_KEYS = ["global_eid", "globalNodeID", "sendProcessID", "sendNeighbourCountCum", "sendNodeIndex"]
STATE_KEYS = ["mass", "boundary0", "dt0", "fi0", "accelerations0", "steps", "Time", "dt", "dt_hist", "displacements",
              "velocities", "accelerations", "boundary", "fi", "f_net"]
INJ_KEYS = ["inj_elems", "inj_gt15", "inj_gt30", "inj_r120", "inj_xsr28", "inj_ps_old", "inj_psxsr", "inj_list95",
            "inj_listx95", "inj_hist95", "inj_histx95", "inj_scalars", "inj_extreme_elems", "inj_volumes",
            "inj_volume_part"]
GP_KEYS = ["detJacobian", "F", "detF", "pk2", "pk2_0", "Eavg", "Hn_1", "Hn_2", "S0n"]


def save(name, dumps, energy, params, keys):
    out = {"nranks": np.int32(len(dumps))}
    for k, v in params.items():
        out["param_" + k] = np.asarray(v)
    if energy is not None:
        out["energy_file"] = energy
    for r, d in enumerate(dumps):
        for k in keys:
            out["r%d_%s" % (r, k)] = d[k]
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f kB)" % (path, os.path.getsize(path) / 1e3))


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="ftmesh_")
    ex = os.path.join(REF, "examples")

    if only == "mixed":
        group_g(tmp)
        shutil.rmtree(tmp, ignore_errors=True)
        return
    if only == "long":
        group_h(tmp)
        shutil.rmtree(tmp, ignore_errors=True)
        return
    if only in ("injury", "rigid"):
        X, conn, pid = mesh.cube_mesh(6, jitter=0.05, nparts_z=3)
        f6 = os.path.join(tmp, "cube6mix.inp")
        mesh.write_abaqus_inp(f6, X, conn, pid)
    else:
        f6 = group_a_to_d(tmp, ex)
    if only != "rigid":
        group_e(tmp, f6)
    if only != "injury":
        group_f(tmp, f6)
    if only == "":
        group_g(tmp)
        group_h(tmp)
    shutil.rmtree(tmp, ignore_errors=True)


def group_a_to_d(tmp, ex):
    # (A) ex9: the reference's only known-answer test (.travis.yml:113-115)
    d, en, _ = run_ref(os.path.join(ex, "ex9", "1-elt-cube.k"), [1], SOFT, 1, 10 ** 9, 1.0, 0.007)
    save("ex9_1elt", d, en[-1:], dict(tMax=1.0, dMax=0.007), MESH_KEYS + STATE_KEYS + GP_KEYS)

    # (B) shipped benchmark mesh, full run, 1/2/4/8 ranks (BASELINE config 1)
    m10 = os.path.join(ex, "Benchmarking-Parallel", "10elements.inp")
    for P in (1, 2, 4, 8):
        d, en, _ = run_ref(m10, [1], SOFT, P, 10 ** 9, 0.1, 0.007)
        keys = MESH_KEYS + MAP_KEYS + STATE_KEYS + (["pk2", "Eavg"] if P == 1 else [])
        save("bench10_p%d" % P, d, en[-1:], dict(tMax=0.1, dMax=0.007), keys)

    # (C) small jittered cubes, every material, 200 steps
    X, conn, pid = mesh.cube_mesh(4, jitter=0.1)
    f4 = os.path.join(tmp, "cube4j.inp")
    mesh.write_abaqus_inp(f4, X, conn, pid)
    HGO = list(BRAIN[:4]) + [10.0, 0, 0, 0, 0]
    cases = {
        "cube4j_m1": (1, SOFT, 0.1), "cube4j_m2": (2, SOFT, 0.1), "cube4j_m3": (3, SOFT, 0.1),
        "cube4j_m4": (4, HGO, 0.004), "cube4j_m5": (5, BRAIN, 0.004),
    }
    for name, (mid, props, tMax) in cases.items():
        d, en, _ = run_ref(f4, [mid], props, 1, 200, tMax, 0.007)
        save(name, d, en[-1:], dict(tMax=tMax, dMax=0.007), MESH_KEYS + STATE_KEYS + GP_KEYS)

    # (D) mixed materials in three z-slabs (parts 0,1,2 = mats 1,4,5), 1 and 3 ranks
    X, conn, pid = mesh.cube_mesh(6, jitter=0.05, nparts_z=3)
    f6 = os.path.join(tmp, "cube6mix.inp")
    mesh.write_abaqus_inp(f6, X, conn, pid)
    STIFF1 = [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0]
    mixprops = STIFF1 + HGO + BRAIN
    for P in (1, 3):
        d, en, _ = run_ref(f6, [1, 4, 5], mixprops, P, 150, 0.004, 0.007)
        save("cube6mix_p%d" % P, d, en[-1:], dict(tMax=0.004, dMax=0.007),
             MESH_KEYS + MAP_KEYS + STATE_KEYS + (GP_KEYS if P == 1 else []))
    return f6


RIGID_TABLES = [  # (t[s], value): angular acceleration x,y,z [rad/s^2], linear acceleration x,y,z [m/s^2]
    ([0.0, 0.002, 0.006], [0.0, 1.5e4, 0.0]), ([0.0, 0.002, 0.006], [0.0, -0.8e4, 0.0]),
    ([0.0, 0.001, 0.003, 0.006], [0.0, 4.0e4, -1.0e4, 0.0]),
    ([0.0, 0.002, 0.006], [0.0, 300.0, 0.0]), ([0.0, 0.002, 0.006], [0.0, 0.0, 0.0]), ([0.0, 0.003, 0.006], [0.0, -200.0, 0.0]),
]


def group_f(tmp, f6):
    # (F) rigid-body prescribed motion of ex5 (ApplyAccBoundaryConditions): the first slab is a rigid part (material 0)
    #     driven by acceleration traces, the others follow; injury criteria on, rigid part excluded
    MED = [1040.0, 2.0e3, 2.0e4, 0, 0, 0, 0, 0, 0]
    HGO_SOFT = [1000.0, 2.0e3, 2.0e4, 500.0, 10.0, 0, 0, 0, 0]
    props = [1500.0, 0, 0, 0, 0, 0, 0, 0, 0] + MED + HGO_SOFT
    d, en, _ = run_ref(f6, [0, 1, 4], props, 1, 400, 0.005, 0.0, injury_exclude=[0], rigid_tables=RIGID_TABLES)
    params = dict(tMax=0.005, dMax=0.0, exclude=[0])
    for k, (t, v) in enumerate(RIGID_TABLES):
        params["rigid_t%d" % k] = t
        params["rigid_v%d" % k] = v
    save("rigid6_p1", d, en[-1:], params,
         MESH_KEYS + ["steps", "Time", "dt", "dt_hist", "displacements", "velocities", "accelerations", "boundary", "fi", "pk2",
                      "rb_y", "rb_ydot", "rb_boundaryID"] + INJ_KEYS)


def group_g(tmp):
    # (G) mixed C3D8 / C3D4 meshes: every third hexahedron of part 0 and all of part 1 split into
    #     six tetrahedra; neo-Hookean + HGO, and neo-Hookean + viscoelastic HGO (history on one-point elements)
    X, conn, pid = mesh.cube_mesh(4, jitter=0.1, nparts_z=2)
    which = (np.arange(conn.shape[0]) % 3 == 1) | (pid == 1)
    flat, eptr, pids, etype = mesh.split_hex_to_tets(conn, pid, which)
    f = os.path.join(tmp, "mix4.inp")
    mesh.write_abaqus_inp_mixed(f, X, flat, eptr, pids, etype)
    MED = [1040.0, 2.0e3, 2.0e4, 0, 0, 0, 0, 0, 0]
    HGO_SOFT = [1000.0, 2.0e3, 2.0e4, 500.0, 10.0, 0, 0, 0, 0]
    VISCO_SOFT = [1000.0, 2.0e3, 2.0e4, 500.0, 10.0, 0.6521, 0.0129, 0.0067, 0.0747]
    for name, mats, props in (("mix4_p1", [1, 4], MED + HGO_SOFT), ("mix4v_p1", [1, 5], MED + VISCO_SOFT)):
        d, en, _ = run_ref(f, mats, props, 1, 150, 0.005, 0.0009, injury_exclude=[])
        save(name, d, en[-1:], dict(tMax=0.005, dMax=0.0009, exclude=[]),
             MESH_KEYS + ["eptr"] + STATE_KEYS + GP_KEYS + INJ_KEYS)


def group_h(tmp):
    # (H) the north-star bar itself: 1000 steps on a jittered mesh for the three materials the kernels specialise
    #     (neo-Hookean, HGO, HGO + Prony), ramp chosen so that the pull reaches ~14 % of the edge after 1000 steps
    X, conn, pid = mesh.cube_mesh(4, jitter=0.1)
    f4 = os.path.join(tmp, "cube4j_long.inp")
    mesh.write_abaqus_inp(f4, X, conn, pid)
    HGO = list(BRAIN[:4]) + [10.0, 0, 0, 0, 0]
    for name, mid, props, tMax, dMax in (("cube4j_m1_1k", 1, SOFT, 4.0, 0.002), ("cube4j_m4_1k", 4, HGO, 0.004, 0.0016),
                                         ("cube4j_m5_1k", 5, BRAIN, 0.004, 0.0016)):
        d, en, _ = run_ref(f4, [mid], props, 1, 1000, tMax, dMax)
        assert int(d[0]["steps"][0]) == 1000, d[0]["steps"]
        save(name, d, en[-1:], dict(tMax=tMax, dMax=dMax), MESH_KEYS + STATE_KEYS + GP_KEYS)


def group_e(tmp, f6):
    # (E) injury criteria of the brain drivers (ex5.cpp:1311-1430) on a large-strain, fast pull: three slabs
    #     (neo-Hookean, HGO, neo-Hookean), the last part excluded from the criteria; 1 and 3 ranks
    MED = [1040.0, 2.0e3, 2.0e4, 0, 0, 0, 0, 0, 0]
    HGO_SOFT = [1000.0, 2.0e3, 2.0e4, 500.0, 10.0, 0, 0, 0, 0]
    injprops = MED + HGO_SOFT + MED
    INJ_SAVE = MESH_KEYS + MAP_KEYS + ["steps", "Time", "dt", "dt_hist", "displacements", "velocities"] + INJ_KEYS
    for name, P, dMax, tMax in (("inj6_p1", 1, 0.0009, 0.005), ("inj6_p3", 3, 0.0009, 0.005), ("inj6b_p1", 1, 0.0012, 0.003)):
        d, en, _ = run_ref(f6, [1, 4, 1], injprops, P, 400, tMax, dMax, injury_exclude=[2])
        save(name, d, en[-1:], dict(tMax=tMax, dMax=dMax, exclude=[2]), INJ_SAVE + (["Eavg", "F"] if P == 1 else []))


if __name__ == "__main__":
    main()
