#!/bin/bash
# final single-GPU pass of the round: full GPU test suite, headline bench line (+ reference arm), ncu launch list and full captures
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --material 4 --no-cpu > gpurun_out/${TAG}_bench_mat4.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 48 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_elem_affine_cj$" -s 4 -c 1 -f -o gpurun_out/${TAG}_k_elem_affine_cj \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_node$" -s 6 -c 3 -f -o gpurun_out/${TAG}_k_node \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_node.log 2>&1
ls gpurun_out | grep ${TAG}
