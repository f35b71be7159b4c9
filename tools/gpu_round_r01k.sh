#!/bin/bash
# Final evidence, part B: the bench lines that go into profiles/.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log10.txt
run() { local name=$1; shift; ( "$@" ) > $O/final_$name.json 2> $O/final_$name.err; echo "== $name: $(python tools/pick.py < $O/final_$name.json) $(tail -1 $O/final_$name.err | cut -c1-200)" | tee -a $O/log10.txt; }
B="python bench.py --steps 100 --warmup 10"
run headline $B
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/final_reference.json 2> $O/final_reference.err; cut -c1-200 $O/final_reference.json | tee -a $O/log10.txt
run jitter $B --no-cpu --jitter 0.05
run general env FTB200_AFFINE=0 $B --no-cpu
run mat4 $B --no-cpu --material 4
run mat5 $B --no-cpu --material 5
run injury $B --no-cpu --injury
