#!/bin/bash
# BASELINE config 3 size on one GPU: 200^3 (8 M elements), neo-Hookean and HGO + Prony.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log11.txt
for mat in 1 5; do
  timeout 75 python bench.py --n 200 --material $mat --steps 20 --warmup 3 --no-cpu > $O/n200_mat$mat.json 2> $O/n200_mat$mat.err
  echo "== n200 mat$mat: $(python tools/pick.py < $O/n200_mat$mat.json) $(tail -1 $O/n200_mat$mat.err | cut -c1-120)" | tee -a $O/log11.txt
done
