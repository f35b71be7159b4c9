#!/bin/bash
# Volume from the Gauss loop + filtered face maximum (dt), injury variant at 8 blocks: full GPU suite and bench lines.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log8.txt
run() { local name=$1; shift; ( "$@" ) > $O/bench_$name.json 2> $O/bench_$name.err; echo "== $name: $(python tools/pick.py < $O/bench_$name.json) $(tail -1 $O/bench_$name.err | cut -c1-200)" | tee -a $O/log8.txt; }
B="python bench.py --steps 100 --warmup 10 --no-cpu"
timeout 500 python -m pytest tests -x -q -m gpu > $O/test8.log 2>&1; echo "rc=$?" >> $O/test8.log
tail -6 $O/test8.log | tee -a $O/log8.txt
run base $B
run jitter $B --jitter 0.05
run general env FTB200_AFFINE=0 $B
run mat4 $B --material 4
run mat5 $B --material 5
run injury $B --injury
