#!/bin/bash
# The reference's own reader + ParMETIS partitioner + femtech_host.cpp under `ftmpirun -np P` on a P-GPU box: one
# ExplicitDynamics() per rank (integration/resident_driver.cpp), throughput line per rank count into gpurun_out/.
#   tools/gpu_dropin_mpi.sh <edge> <tMax> <np list...>
n=${1:-40}; tmax=${2:-0.01}; shift 2
work=$(mktemp -d); root=$PWD
python - "$n" "$work" <<'PY'
import sys
sys.path.insert(0, ".")
from femtech_b200 import mesh
n, work = int(sys.argv[1]), sys.argv[2]
X, conn, pid = mesh.cube_mesh(n)
mesh.write_abaqus_inp(work + "/cube.inp", X, conn, pid)
mesh.write_materials_dat(work + "/materials.dat", [1], [1040.0, 2.0e5, 4.0e5, 0, 0, 0, 0, 0, 0])
PY
cd $work
for p in "$@"; do
  if [ "$p" = 1 ]; then cmd="$root/oracle/_ref/dropin_resident"; else cmd="$root/oracle/_ref/ftmpirun -np $p $root/oracle/_ref/dropin_resident"; fi
  CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 $cmd cube.inp $tmax 0.0007 0.005 2>&1 | grep "^RESIDENT" | tee -a $root/gpurun_out/r02_dropin_mpi_n$n.txt
done
