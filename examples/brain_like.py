#!/usr/bin/env python
"""A brain-simulation-shaped run on one GPU (BASELINE config 5 in miniature): a multi-part hex8 box -- rigid outer shell
(material 0, "skull"), a soft neo-Hookean layer ("CSF"), a viscoelastic HGO core ("brain", the properties of
examples/ex5/materials.dat) -- driven by a prescribed rigid-body motion of the shell (rotational + linear acceleration
pulses, ex5.cpp:339-371), with the injury criteria evaluated every step (ex5.cpp:1311-1430) and a ParaView file at the end.

    python examples/brain_like.py --n 100 --t-end 0.004 [--vtu out.vtu]

Everything runs in the resident loop: no per-step host work besides the optional progress poll."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femtech_b200 import io as fio  # noqa: E402
from femtech_b200 import mesh, solver  # noqa: E402

BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]  # examples/ex5/materials.dat:2


def build(n, L=0.16):
    X, conn, _ = mesh.cube_mesh(n, L=L)
    X = X - 0.5 * L  # rotation about the centre, like TransformMesh (ex5.cpp:1432-1500) centres the head
    idx = np.arange(n ** 3)
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    depth = np.minimum.reduce([i, j, k, n - 1 - i, n - 1 - j, n - 1 - k])
    pid = np.where(depth == 0, 0, np.where(depth == 1, 1, 2)).astype(np.int32)
    return X, conn, pid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=60)
    ap.add_argument("--t-end", type=float, default=0.002)
    ap.add_argument("--max-steps", type=int, default=10 ** 9)
    ap.add_argument("--vtu", default="")
    args = ap.parse_args()
    X, conn, pid = build(args.n)
    # parts must first appear in ascending order for the reference's reader; here they go straight to the C-ABI
    props = [1500.0, 0, 0, 0, 0, 0, 0, 0, 0] + [1040.0, 1.0e4, 2.0e8, 0, 0, 0, 0, 0, 0] + BRAIN
    m = solver.FemTech(X, conn, pid, [0, 1, 5], props)
    m.ShapeFunctions()
    m.AssembleLumpedMass()
    tp = 0.4 * args.t_end
    tables = [([0.0, tp, args.t_end], [0.0, 4.0e3, 0.0]), ([0.0, tp, args.t_end], [0.0, -2.0e3, 0.0]),
              ([0.0, tp, args.t_end], [0.0, 6.0e3, 0.0]),
              ([0.0, tp, args.t_end], [0.0, 50.0 * 9.81, 0.0]), ([0.0, args.t_end], [0.0, 0.0]), ([0.0, args.t_end], [0.0, 0.0])]
    m.set_rigid_bc(tables)
    m.FailureTimeStep = 1e-11
    m.explicit_begin(energy_every=1)
    m.InitInjuryCriterion(exclude_pids=[0, 1])
    t0 = time.perf_counter()
    steps = m.ExplicitDynamics(args.t_end, maxSteps=args.max_steps)
    wall = time.perf_counter() - t0
    r = m.injury_results()
    e = m.energy()
    y, _, nb = m.rigid_state()
    nE = conn.shape[0]
    print("elements %d (rigid %d, soft %d, viscoelastic %d), rigid-motion nodes %d" % (nE, (pid == 0).sum(), (pid == 1).sum(), (pid == 2).sum(), nb))
    print("steps %d to t = %.4e s, dt = %.3e, %.3f s wall -> %.3e element-steps/s (injury criteria on)" % (steps, m.Time, m.dt, wall, nE * steps / wall))
    print("shell: omega = (%.2f %.2f %.2f) rad/s, displacement = (%.2e %.2e %.2e) m" % (*y[0:3], *y[9:12]))
    print("energy: Wint %.4e Wext %.4e WKE %.4e |balance| %.2e" % tuple(e))
    print("MPS max %.4f (element %d, t = %.3e), MPS-95 %.4f, MPSxSR-95 %.3f, CSDM-15 volume fraction %.4f" %
          (r["scalars"][0], r["extreme_elems"][0], r["scalars"][1], r["scalars"][8], r["scalars"][10], r["volumes"][0] / max(r["volumes"][4], 1e-300)))
    if args.vtu:
        fl = r["flags"]
        fio.WriteVTU(args.vtu, X.reshape(-1), m.displacements, conn.reshape(-1), 8 * np.arange(nE + 1), ["C3D8"] * nE, pid,
                     accelerations=m.accelerations, boundary=m.boundary, Eavg=m.CalculateStrain(),
                     int_cell_data={"CSDM-15": fl & 1, "CSDM-30": (fl >> 1) & 1, "PSR-120": (fl >> 2) & 1, "PSxSR-28": (fl >> 3) & 1,
                                    "MPS-95": (fl >> 4) & 1})
        print("wrote", args.vtu)
    m.close()
    return steps, r, e


if __name__ == "__main__":
    main()
