#!/bin/bash
# Final evidence, part A: full GPU suite, ncu full of both element kernels and of k_node, launch list.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log9.txt
timeout 500 python -m pytest tests -q -m gpu > $O/test9.log 2>&1; echo "rc=$?" >> $O/test9.log
tail -6 $O/test9.log | tee -a $O/log9.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_elem_affine --launch-skip 20 --launch-count 1 \
  -f -o $O/k_elem_affine python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_a.log 2>&1; echo "ncu affine rc=$?" | tee -a $O/log9.txt
FTB200_AFFINE=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_elem<" --launch-skip 20 --launch-count 1 \
  -f -o $O/k_elem_general python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_g.log 2>&1; echo "ncu general rc=$?" | tee -a $O/log9.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $O/log9.txt
