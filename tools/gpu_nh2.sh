#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_affine.py tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -x -q -k "peer or p2p or affine or two_gpu or current_jacobian or injury" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench1.json')); print('N=1 value %.4e ms/step %.4f elem %.4f node %.4f part-loop %s' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['k_node']['launch_ms'], (d.get('n1_partitioned_loop') or {}).get('ms_per_step')))"
