"""brain-like config (rigid shell, exclusions, mixed materials) on P in-process ranks vs one context: percentile histories"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import numpy as np
from femtech_b200 import mesh, solver, dist as fdist
import brain_like_dist as B
P = int(sys.argv[1]); n = int(sys.argv[2]); nsteps = int(sys.argv[3]); p2p = len(sys.argv) > 4 and sys.argv[4] == "p2p"
L, t_end = 0.16, 0.002
pg = fdist.proc_grid(P)
loc = (n // pg[0], n // pg[1], n // pg[2])
parts = []
for r in range(P):
    p = fdist.brick_partition(loc, pg, r, L_local=L * loc[0] / n)
    p["coordinates"] = p["coordinates"] - 0.5 * L
    rx, ry, rz = r % pg[0], (r // pg[0]) % pg[1], r // (pg[0] * pg[1])
    e = np.arange(loc[0] * loc[1] * loc[2])
    p["pid"] = B.part_ids(e % loc[0] + rx * loc[0], (e // loc[0]) % loc[1] + ry * loc[1], e // (loc[0] * loc[1]) + rz * loc[2], n)
    parts.append(p)
X, conn, _ = mesh.box_mesh(n, n, n, L / n)
ee = np.arange(n ** 3)
s = solver.FemTech(X - 0.5 * L, conn, B.part_ids(ee % n, (ee // n) % n, ee // (n * n), n), B.MATS, B.PROPS)
s.ShapeFunctions(); s.AssembleLumpedMass(); s.set_rigid_bc(B.tables(t_end))
s._check(s.L.ftb200_record_history(s._h, nsteps + 8))
s.explicit_begin(energy_every=1); s.InitInjuryCriterion(exclude_pids=[0, 1])
assert s.ExplicitDynamics(t_end, maxSteps=nsteps) == nsteps
g95, gx95 = s.injury_history(0, nsteps); U = s.displacements.reshape(-1, 3).copy(); s.close()
grp = fdist.LocalGroup(parts, B.MATS, B.PROPS)
grp.setup()
for m in grp.models:
    m.set_rigid_bc(B.tables(t_end)); m._check(m.L.ftb200_record_history(m._h, nsteps + 8))
grp.explicit_begin(energy_every=1)
grp.InitInjuryCriterion(exclude_pids=[0, 1])
if p2p:
    grp.enable_p2p(); grp.run_p2p(t_end, nsteps)
else:
    grp.run(t_end, nsteps)
for r, (m, p) in enumerate(zip(grp.models, parts)):
    m.sync_out()
    h95, hx95 = m.injury_history(0, nsteps)
    bad = np.nonzero(np.abs(h95 - g95) > 1e-9 * np.abs(g95).max())[0]
    print("rank", r, "u err", float(np.abs(m.displacements.reshape(-1, 3) - U[p["node_gids"]]).max() / np.abs(U).max()), "h95 bad steps", bad.size, "first", (int(bad[0]) if bad.size else None),
          [(float(h95[i]), float(g95[i])) for i in bad[:2]])
grp.close()
