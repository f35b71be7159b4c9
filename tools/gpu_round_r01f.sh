#!/bin/bash
# Round 1, GPU call 6 of the second session: resident driver through the reference's API, graph kept across explicit_begin,
# full GPU suite, default bench line.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log5.txt
echo "== full gpu suite" | tee -a $O/log5.txt
timeout 500 python -m pytest tests -x -q -m gpu > $O/test5.log 2>&1; echo "rc=$?" >> $O/test5.log
tail -12 $O/test5.log | tee -a $O/log5.txt
echo "== bench default" | tee -a $O/log5.txt
timeout 400 python bench.py --steps 100 --warmup 10 > $O/bench_r01_final.json 2> $O/bench_r01_final.err; echo "rc=$?" | tee -a $O/log5.txt
python tools/pick.py < $O/bench_r01_final.json | tee -a $O/log5.txt
tail -3 $O/bench_r01_final.err | tee -a $O/log5.txt
