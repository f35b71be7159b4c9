"""single-partition loop vs partitioned (peer-memory) loop on one GPU with the rigid-body BC: where do they part?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
import numpy as np
from femtech_b200 import mesh, solver
import brain_like_dist as B
n = int(sys.argv[1]); steps = int(sys.argv[2]); chunks = [int(c) for c in sys.argv[3].split(",")]
L = 0.16
X, conn, _ = mesh.box_mesh(n, n, n, L / n)
ee = np.arange(n ** 3)
pid = B.part_ids(ee % n, (ee // n) % n, ee // (n * n), n)
MODE = os.environ.get("PROBE_MODE", "rigid")
def run(part, chunks):
    if MODE == "rigid":
        s = solver.FemTech(X - 0.5 * L, conn, pid, B.MATS, B.PROPS)
    elif MODE == "bc135":   # same parts, no rigid material: 1 / 1 / 5, benchmark BC
        s = solver.FemTech(X, conn, pid, [1, 1, 5], B.PROPS[9:18] + B.PROPS[9:18] + B.PROPS[18:27])
    elif MODE == "bc5":
        s = solver.FemTech(X, conn, 0 * pid, [5], B.PROPS[18:27])
    else:
        s = solver.FemTech(X, conn, 0 * pid, [1], B.PROPS[9:18])
    s.ShapeFunctions(); s.AssembleLumpedMass()
    if MODE == "rigid":
        s.set_rigid_bc(B.tables(0.002))
    else:
        kind, rate = mesh.benchmark_bc(X, L=L, dMax=1.75, tMax=1.0)
        s.set_bc(kind, rate)
    s._check(s.L.ftb200_record_history(s._h, steps + 8))
    s.explicit_begin(energy_every=1)
    if part: s.enable_partitioned_loop()
    for c in chunks:
        s.run_async(0.002, c)
    s._poll(); s.sync_out(forces=False)
    dth, eh = s.history(0, steps)
    y = s.rigid_state()[0] if MODE == 'rigid' else np.zeros(12)
    out = (s.displacements.copy(), dth, s.Time, y.copy(), int(s.steps_done), eh, int(s.status_bits))
    s.close()
    return out
a = run(False, [steps]); b = run(True, chunks); c = run(False, chunks)
for name, r in (("partitioned", b), ("single chunked", c)):
    k = np.nonzero(r[1] != a[1])[0]
    print(name, "steps", r[4], "Time", r[2], a[2], "first dt diff at", (int(k[0]) if k.size else None), "max u diff", float(np.abs(r[0] - a[0]).max()), "umax", float(np.abs(a[0]).max()), "y diff", float(np.abs(r[3] - a[3]).max()))

if os.environ.get("PROBE_DUMP"):
    np.set_printoptions(precision=17, linewidth=200)
    for i in range(0, min(steps, 6)):
        print(i, "dt a/c", repr(a[1][i]), repr(c[1][i]), "E a", a[5][i], "E c", c[5][i])
    print("status", a[6], c[6]); print("y a", a[3]); print("y c", c[3])
    d = np.abs(a[0] - c[0]).reshape(-1, 3).max(axis=1)
    k = np.argsort(-d)[:5]
    print("worst nodes", k, d[k], (X - 0.5 * L)[k])

if os.environ.get("PROBE_DUMP") and MODE == "rigid":
    y = a[3]
    r = y[3:6]; mag = np.linalg.norm(r)
    q = np.array([np.cos(mag), *(np.sin(mag) / mag * r)]) if mag > 0 else np.array([1.0, 0, 0, 0])
    def qmul(p, q_):
        return np.array([p[0]*q_[0]-p[1]*q_[1]-p[2]*q_[2]-p[3]*q_[3], p[0]*q_[1]+p[1]*q_[0]+p[2]*q_[3]-p[3]*q_[2],
                         p[0]*q_[2]-p[1]*q_[3]+p[2]*q_[0]+p[3]*q_[1], p[0]*q_[3]+p[1]*q_[2]-p[2]*q_[1]+p[3]*q_[0]])
    qi = np.array([q[0], -q[1], -q[2], -q[3]]) / (q @ q)
    Xc = X - 0.5 * L
    for node in k[:3]:
        V = np.array([0.0, *Xc[node]])
        up = qmul(qmul(q, V), qi)[1:] - Xc[node] + y[9:12]
        print("node", node, "analytic", up, "a", a[0].reshape(-1, 3)[node], "c", c[0].reshape(-1, 3)[node])
