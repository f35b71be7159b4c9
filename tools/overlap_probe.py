"""Tuning probe: can k_node (HBM bound) run concurrently with k_elem (fp64 bound)?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femtech_b200 import mesh, solver
X, conn, pid = mesh.cube_mesh(100)
m = solver.FemTech(X, conn, pid, [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])
m.ShapeFunctions(); m.AssembleLumpedMass()
kind, rate = mesh.benchmark_bc(X); m.set_bc(kind, rate); m.explicit_begin(energy_every=0)
m.run_async(1e30, 10); m._poll()
f = m.L.ftb200_debug_overlap
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
for conc in (0, 1, 0, 1):
    out = C.c_double()
    rc = f(m._h, 20, conc, C.byref(out))
    print("concurrent=%d rc=%d ms per (k_elem + k_node) = %.4f" % (conc, rc, out.value))
