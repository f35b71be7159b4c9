#!/bin/bash
# 2-GPU check: the skipped 2-GPU tests, then the scaling bench line at N = 2 (peer-memory transport) and its reference arm.
set -u
O=gpurun_out
mkdir -p $O
: > $O/log6.txt
timeout 300 python -m pytest tests -x -q -m gpu -k "two_gpu or peer_memory or multirank" > $O/test6.log 2>&1; echo "rc=$?" >> $O/test6.log
tail -5 $O/test6.log | tee -a $O/log6.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_scale2.json 2> $O/bench_scale2.err; echo "rc=$?" | tee -a $O/log6.txt
grep '^{' $O/bench_scale2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value=%.3e ms/step=%.4f e2e=%.3e valid=%s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('valid')))" | tee -a $O/log6.txt
tail -3 $O/bench_scale2.err | tee -a $O/log6.txt
