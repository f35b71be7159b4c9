#!/bin/bash
# N GPUs: the partitioned step under each FTB200_P2P_ORDER mode, traced (run with gpurun --gpus N)
TAG=${1:-r02r}
N=${2:-2}
MODES=${3:-"1 2 3"}
mkdir -p gpurun_out
port=29520
for m in $MODES; do
  port=$((port+1))
  export FTB200_P2P_ORDER=$m
  export FTB200_P2P_TRACE=$PWD/gpurun_out/${TAG}_order${m}_trace
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 5 --no-cpu --no-validate > gpurun_out/${TAG}_order${m}.json 2> gpurun_out/${TAG}_order${m}.err
  echo "== FTB200_P2P_ORDER=$m"
  python -c "
import json,sys
for l in reversed(open('gpurun_out/${TAG}_order${m}.json').read().splitlines()):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N value %.4e ms/step %.4f valid %s' % (d['value'], d['ms_per_step'], d['valid'])); break
else: print(open('gpurun_out/${TAG}_order${m}.err').read()[-600:])" 2>&1 | tail -4
  python tools/p2p_trace_report.py gpurun_out/${TAG}_order${m}_trace 10 2>&1 | head -11
done
