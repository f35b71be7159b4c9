#!/bin/bash
# Round 1, GPU call 4 of the second session: the overlapped step -- bitwise test against the serial step, then bench sweeps.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log4.txt
run() { local name=$1; shift; ( "$@" ) > $O/bench_$name.json 2> $O/bench_$name.err; echo "== $name: $(python tools/pick.py < $O/bench_$name.json) $(tail -1 $O/bench_$name.err | cut -c1-200)" | tee -a $O/log4.txt; }
B="python bench.py --steps 100 --warmup 10 --no-cpu"
echo "== overlap tests" | tee -a $O/log4.txt
timeout 300 python -m pytest tests/test_gpu_overlap.py -x -q -m gpu > $O/test4.log 2>&1; echo "rc=$?" >> $O/test4.log
tail -15 $O/test4.log | tee -a $O/log4.txt
run serial env FTB200_OVERLAP=0 $B
run ovl env FTB200_OVERLAP=1 $B
run ovl_e6 env FTB200_OVERLAP=1 FTB200_OVL_ELEM_BLOCKS=6 $B
run ovl_e6n2 env FTB200_OVERLAP=1 FTB200_OVL_ELEM_BLOCKS=6 FTB200_OVL_NODE_BLOCKS=2 $B
run ovl_e5n3 env FTB200_OVERLAP=1 FTB200_OVL_ELEM_BLOCKS=5 FTB200_OVL_NODE_BLOCKS=3 $B
run ovl_jit env FTB200_OVERLAP=1 $B --jitter 0.05
run ovl_mat4 env FTB200_OVERLAP=1 $B --material 4
echo "== ncu launch list (overlap)" | tee -a $O/log4.txt
FTB200_OVERLAP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_ovl.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_list_ovl.log 2>&1; echo "rc=$?" | tee -a $O/log4.txt
