#!/bin/bash
# Programmatic dependent launch of the step's kernels: full GPU suite with it on, then bench on / off.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log12.txt
FTB200_PDL=1 timeout 80 python -m pytest tests -q -m gpu -x > $O/test12.log 2>&1; echo "rc=$?" >> $O/test12.log
tail -5 $O/test12.log | tee -a $O/log12.txt
run() { local name=$1; shift; ( "$@" ) > $O/bench_$name.json 2> $O/bench_$name.err; echo "== $name: $(python tools/pick.py < $O/bench_$name.json) $(tail -1 $O/bench_$name.err | cut -c1-160)" | tee -a $O/log12.txt; }
run pdl1 env FTB200_PDL=1 timeout 40 python bench.py --steps 100 --warmup 10 --no-cpu
run pdl0 env FTB200_PDL=0 timeout 40 python bench.py --steps 100 --warmup 10 --no-cpu
