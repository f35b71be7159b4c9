#!/bin/bash
# One gpurun call of round 1 (second session): parity of the parallelepiped element kernel, bench lines of the
# variants, ncu launch list and one full capture of k_elem_affine.  Everything lands in gpurun_out/.
# Steps are ordered by priority and individually bounded so that a slow one cannot eat the budget.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > $O/gpu.txt 2>&1
echo "== affine parity tests" | tee $O/log.txt
timeout 420 python -m pytest tests/test_gpu_affine.py -x -q -m gpu > $O/test_affine.log 2>&1; echo "rc=$?" >> $O/test_affine.log
tail -3 $O/test_affine.log | tee -a $O/log.txt
echo "== bench default" | tee -a $O/log.txt
timeout 400 python bench.py --steps 100 --warmup 10 > $O/bench_affine.json 2> $O/bench_affine.err; echo "rc=$?" | tee -a $O/log.txt
python tools/pick.py < $O/bench_affine.json | tee -a $O/log.txt
echo "== bench general kernel on the structured cube (FTB200_AFFINE=0)" | tee -a $O/log.txt
FTB200_AFFINE=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $O/bench_general.json 2> $O/bench_general.err
python tools/pick.py < $O/bench_general.json | tee -a $O/log.txt
for v in mb6 mb6h; do
  echo "== bench lib $v" | tee -a $O/log.txt
  FTB200_LIB=$PWD/femtech_b200/libftb200_$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $O/bench_$v.json 2> $O/bench_$v.err
  python tools/pick.py < $O/bench_$v.json | tee -a $O/log.txt
done
echo "== bench jittered cube" | tee -a $O/log.txt
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --jitter 0.05 > $O/bench_jitter.json 2> $O/bench_jitter.err
python tools/pick.py < $O/bench_jitter.json | tee -a $O/log.txt
echo "== ncu full capture of k_elem_affine" | tee -a $O/log.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_elem_affine --launch-skip 20 --launch-count 1 \
  -f -o $O/k_elem_affine python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1; echo "rc=$?" | tee -a $O/log.txt
echo "== ncu launch list" | tee -a $O/log.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_list.log 2>&1; echo "rc=$?" | tee -a $O/log.txt
for mat in 4 5; do
  echo "== bench material $mat" | tee -a $O/log.txt
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --material $mat > $O/bench_mat$mat.json 2> $O/bench_mat$mat.err
  python tools/pick.py < $O/bench_mat$mat.json | tee -a $O/log.txt
done
echo "== full gpu suite" | tee -a $O/log.txt
timeout 600 python -m pytest tests -x -q -m gpu > $O/test_gpu.log 2>&1; echo "rc=$?" >> $O/test_gpu.log
tail -3 $O/test_gpu.log | tee -a $O/log.txt
