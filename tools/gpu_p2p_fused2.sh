#!/bin/bash
TAG=${1:-r02w}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_dropin_driver.py -m gpu -x -q -k "two_gpu or peer_memory" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu_p2p_fused.sh $TAG $N
