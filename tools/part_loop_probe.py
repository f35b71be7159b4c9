"""N = 1 through the partitioned (peer-memory) loop, for an ncu launch list (single process):
   ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file out.csv python tools/part_loop_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from femtech_b200 import mesh, solver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
X, conn, pid = mesh.cube_mesh(n)
kind, rate = mesh.benchmark_bc(X, dMax=0.07, tMax=1.0)
m = solver.FemTech(X, conn, pid, [1], [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0])
m.ShapeFunctions(); m.AssembleLumpedMass(); m.set_bc(kind, rate)
m.enable_partitioned_loop()
m.explicit_begin(energy_every=1)
m.profile(True)   # direct launches: every kernel is visible to the profiler by name
m.run_async(1e30, 12)
m._poll()
print("steps", m.steps_done, "Time", m.Time)
