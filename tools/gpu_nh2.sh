#!/bin/bash
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_affine.py tests/test_gpu_step_ring.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu_variants.sh > gpurun_out/${TAG}_variants.txt 2>&1; cat gpurun_out/${TAG}_variants.txt
