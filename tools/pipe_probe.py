"""Tuning probe: standalone time of the two persistent pipe kernels (dependencies pre-satisfied)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femtech_b200 import mesh, solver
X, conn, pid = mesh.cube_mesh(100)
m = solver.FemTech(X, conn, pid, [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])
m.ShapeFunctions(); m.AssembleLumpedMass()
kind, rate = mesh.benchmark_bc(X); m.set_bc(kind, rate); m.explicit_begin(energy_every=int(os.environ.get("ENERGY", "1")))
f = m.L.ftb200_debug_pipe
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
for which in (0, 1, 0, 1):
    out = C.c_double()
    rc = f(m._h, which, 5, C.byref(out))
    print("which=%s rc=%d ms per launch = %.4f" % ("elem" if which == 0 else "node", rc, out.value))
