#!/bin/bash
# 2 GPUs: the partitioned step with and without the diagnostic trace, 100^3 per GPU (run with gpurun --gpus 2)
TAG=${1:-r02q}
N=${2:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu --no-validate; }
timeout 300 bash -c "$(declare -f run); N=$N; run 29511" > gpurun_out/${TAG}_scale${N}.json 2> gpurun_out/${TAG}_scale${N}.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_scale${N}.json')); print('N=$N value %.4e ms/step %.4f valid %s' % (d['value'], d['ms_per_step'], d['valid']))"
export FTB200_P2P_TRACE=$PWD/gpurun_out/${TAG}_trace
timeout 300 bash -c "$(declare -f run); N=$N; run 29512" > gpurun_out/${TAG}_scale${N}_traced.json 2> gpurun_out/${TAG}_scale${N}_traced.err
unset FTB200_P2P_TRACE
python tools/p2p_trace_report.py gpurun_out/${TAG}_trace 10 | tee gpurun_out/${TAG}_trace_report.txt
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench1.json')); print('N=1 value %.4e ms/step %.4f part-loop %s' % (d['value'], d['ms_per_step'], (d.get('n1_partitioned_loop') or {}).get('ms_per_step')))"
