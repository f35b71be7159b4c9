#!/bin/bash
# Round 1, GPU call 3 of the second session: step ring, k_node 128x8, full GPU suite, bench lines, ncu launch list + k_node.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log3.txt
echo "== full gpu suite" | tee -a $O/log3.txt
timeout 420 python -m pytest tests -x -q -m gpu > $O/test3.log 2>&1; echo "rc=$?" >> $O/test3.log
tail -4 $O/test3.log | tee -a $O/log3.txt
echo "== bench default" | tee -a $O/log3.txt
timeout 400 python bench.py --steps 100 --warmup 10 > $O/bench_r01_final.json 2> $O/bench_r01_final.err; echo "rc=$?" | tee -a $O/log3.txt
python tools/pick.py < $O/bench_r01_final.json | tee -a $O/log3.txt
tail -3 $O/bench_r01_final.err | tee -a $O/log3.txt
echo "== reference arm" | tee -a $O/log3.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" | tee -a $O/log3.txt
cut -c1-300 $O/bench_reference.json | tee -a $O/log3.txt
echo "== ncu launch list" | tee -a $O/log3.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_list.log 2>&1; echo "rc=$?" | tee -a $O/log3.txt
echo "== ncu full k_node" | tee -a $O/log3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_node --launch-skip 12 --launch-count 1 \
  -f -o $O/k_node python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_node.log 2>&1; echo "rc=$?" | tee -a $O/log3.txt
for a in "--jitter 0.05" "--material 4" "--material 5" "--injury"; do
  name=$(echo $a | tr -d ' -.')
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu $a > $O/bench_$name.json 2> $O/bench_$name.err
  echo "== $a: $(python tools/pick.py < $O/bench_$name.json)" | tee -a $O/log3.txt
done
