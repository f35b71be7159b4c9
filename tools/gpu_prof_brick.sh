#!/bin/bash
# ncu captures of the brick-fused step at 100^3 (run under gpurun): launch list + full capture of k_brick and k_surf
set -x
export FTB200_BRICK=1
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_brick$" -s 4 -c 1 -f -o gpurun_out/${TAG}_k_brick \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_brick.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_surf -s 4 -c 1 -f -o gpurun_out/${TAG}_k_surf \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_surf.log 2>&1
ls -la gpurun_out/
