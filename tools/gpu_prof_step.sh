#!/bin/bash
# ncu evidence for the default (two-kernel) step at 100^3 (run under gpurun, one GPU): launch list + full captures of the
# element and node kernels.  TAG names the files under gpurun_out/.
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 48 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_elem_affine_cj$" -s 4 -c 1 -f -o gpurun_out/${TAG}_k_elem_affine \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_node$" -s 5 -c 1 -f -o gpurun_out/${TAG}_k_node \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_node.log 2>&1
ls -la gpurun_out/ | tail -8
