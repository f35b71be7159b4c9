#!/bin/bash
# Gauss-loop unroll factors of the affine kernel; resident blocks of its injury variant.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log7.txt
run() { local name=$1; shift; ( "$@" ) > $O/bench_$name.json 2> $O/bench_$name.err; echo "== $name: $(python tools/pick.py < $O/bench_$name.json) $(tail -1 $O/bench_$name.err | cut -c1-200)" | tee -a $O/log7.txt; }
B="python bench.py --steps 100 --warmup 10 --no-cpu"
run base $B
for v in u2 u8; do run $v env FTB200_LIB=$PWD/femtech_b200/libftb200_$v.so $B; done
run u8_mat4 env FTB200_LIB=$PWD/femtech_b200/libftb200_u8.so $B --material 4
run inj_base $B --injury
for v in inj6 inj8; do run $v env FTB200_LIB=$PWD/femtech_b200/libftb200_$v.so $B --injury; done
