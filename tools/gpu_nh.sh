#!/bin/bash
# round 2, current-Jacobian neo-Hookean kernel: parity on the GPU, then the unroll variants and the old kernel on the same box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_affine.py -m gpu -x -q > gpurun_out/r02m_pytest_affine.log 2>&1; tail -3 gpurun_out/r02m_pytest_affine.log
bash tools/gpu_variants.sh > gpurun_out/r02m_variants.txt 2>&1
FTB200_NH=0 FTB200_LIB=$PWD/femtech_b200/libftb200.so timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02m_bench_old_kernel.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_old_kernel.json')); print('old kernel (FTB200_NH=0): value %.4e ms/step %.4f elem %.4f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms']))" >> gpurun_out/r02m_variants.txt
cat gpurun_out/r02m_variants.txt
timeout 60 tools/issue_mix > gpurun_out/r02m_issue_mix.txt 2>&1; cat gpurun_out/r02m_issue_mix.txt
