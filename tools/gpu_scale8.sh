#!/bin/bash
TAG=${1:-r02y}
N=${2:-8}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > gpurun_out/${TAG}_scale${N}.json 2> gpurun_out/${TAG}_scale${N}.err
python -c "
import json,sys
for l in reversed(open(sys.argv[1]).read().splitlines()):
    if l.startswith('{'):
        d=json.loads(l); print('N=%d value %.4e ms/step %.4f valid %s validation %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['valid'], (d.get('validation') or {}))); break
else: print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
" gpurun_out/${TAG}_scale${N}.json
