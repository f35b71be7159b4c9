#!/bin/bash
# round 2, current-Jacobian kernels: parity on the GPU, then library variants (experiments/_lib: material 1,
# experiments/_lib4: material 4) and the displacement-gradient kernels (FTB200_NH=0) on the same box
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_affine.py -m gpu -x -q > gpurun_out/${TAG}_pytest_affine.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_affine.log
bash tools/gpu_variants.sh > gpurun_out/${TAG}_variants.txt 2>&1
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print('%-40s value %.4e ms/step %.4f elem %.4f node %.4f' % (sys.argv[2], d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline'].get('k_node',{}).get('launch_ms',0)))" "$1" "$2" >> gpurun_out/${TAG}_variants.txt; }
FTB200_NH=0 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu > /tmp/o.json 2>/dev/null; line /tmp/o.json "mat 1 FTB200_NH=0"
for f in experiments/_lib4/lib_*.so; do
  FTB200_LIB=$PWD/$f timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu --material 4 > /tmp/o.json 2>/dev/null; line /tmp/o.json "mat 4 $f"
done
FTB200_NH=0 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu --material 4 > /tmp/o.json 2>/dev/null; line /tmp/o.json "mat 4 FTB200_NH=0"
cat gpurun_out/${TAG}_variants.txt
