#!/bin/bash
# N GPUs: the partitioned step with the exchange fused into the boundary element kernel: traced run, then the
# validated bench line (state check against the single-GPU run of the global mesh)
TAG=${1:-r02t}
N=${2:-2}
mkdir -p gpurun_out
lastjson() { python -c "
import json,sys
for l in reversed(open(sys.argv[1]).read().splitlines()):
    if l.startswith('{'):
        d=json.loads(l); print('%s: N=%d value %.4e ms/step %.4f valid %s validation %s' % (sys.argv[2], d['n_gpus'], d['value'], d['ms_per_step'], d['valid'], (d.get('validation') or {}))); break
" "$1" "$2"; }
FTB200_P2P_TRACE=$PWD/gpurun_out/${TAG}_trace timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu --no-validate > gpurun_out/${TAG}_traced.json 2> gpurun_out/${TAG}_traced.err
lastjson gpurun_out/${TAG}_traced.json traced
python tools/p2p_trace_report.py gpurun_out/${TAG}_trace 10 2>&1 | head -11 | tee gpurun_out/${TAG}_trace_report.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > gpurun_out/${TAG}_scale${N}.json 2> gpurun_out/${TAG}_scale${N}.err
lastjson gpurun_out/${TAG}_scale${N}.json validated
tail -3 gpurun_out/${TAG}_scale${N}.err | cut -c1-300
