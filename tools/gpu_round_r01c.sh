#!/bin/bash
# Round 1, GPU call 2 of the second session: L2 prefetch distance sweep of the element kernels, k_node launch-shape
# variants, parity tests with the prefetch on.  Results in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log2.txt
run() {  # name, command...
  local name=$1; shift
  ( "$@" ) > $O/bench_$name.json 2> $O/bench_$name.err
  echo "== $name: $(python tools/pick.py < $O/bench_$name.json)" | tee -a $O/log2.txt
}
B="python bench.py --steps 100 --warmup 10 --no-cpu"
echo "== parity tests (prefetch on, default)" | tee -a $O/log2.txt
timeout 300 python -m pytest tests/test_gpu_affine.py tests/test_gpu_parity.py -x -q -m gpu > $O/test2.log 2>&1; echo "rc=$?" >> $O/test2.log
tail -3 $O/test2.log | tee -a $O/log2.txt
for w in 0 1 2 3 4; do run pf$w env FTB200_PREFETCH_WAVES=$w $B; done
for w in 0 2; do run jit_pf$w env FTB200_PREFETCH_WAVES=$w $B --jitter 0.05; done
for v in nb128 nm4 nb128m8; do run $v env FTB200_LIB=$PWD/femtech_b200/libftb200_$v.so $B; done
run mat4 $B --material 4
run mat5 $B --material 5
