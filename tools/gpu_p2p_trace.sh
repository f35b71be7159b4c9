#!/bin/bash
# Phase breakdown of the partitioned step on N GPUs (run under `gpurun --gpus N`), 100^3 per GPU:
#   tools/gpu_p2p_trace.sh <tag> <N> ["<FTB200_P2P_ORDER modes>"] [FTB200_P2P_FUSED]
# For every mode (default "0"): one traced bench run (FTB200_P2P_TRACE, tools/p2p_trace_report.py); then the
# validated bench line (state check against the single-GPU run of the global mesh) with the first mode.
TAG=${1:-r02}
N=${2:-2}
MODES=${3:-"0"}
export FTB200_P2P_FUSED=${4:-1}
mkdir -p gpurun_out
port=29520
line() { python -c "
import json, sys
for l in reversed(open(sys.argv[1]).read().splitlines()):
    if l.startswith('{'):
        d = json.loads(l); print('%s: N=%d value %.4e ms/step %.4f valid %s %s' % (sys.argv[2], d['n_gpus'], d['value'], d['ms_per_step'], d['valid'], d.get('validation') or '')); break
else: print(sys.argv[2], 'FAILED', open(sys.argv[1].replace('.json', '.err')).read()[-600:])" "$1" "$2"; }
run() { port=$((port + 1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
          bench.py --gpus $N --steps 100 --warmup 5 --no-cpu "$@"; }
for m in $MODES; do
  export FTB200_P2P_ORDER=$m
  FTB200_P2P_TRACE=$PWD/gpurun_out/${TAG}_order${m}_trace run --no-validate > gpurun_out/${TAG}_order${m}_traced.json 2> gpurun_out/${TAG}_order${m}_traced.err
  line gpurun_out/${TAG}_order${m}_traced.json "traced, FTB200_P2P_ORDER=$m FTB200_P2P_FUSED=$FTB200_P2P_FUSED"
  python tools/p2p_trace_report.py gpurun_out/${TAG}_order${m}_trace 10 > gpurun_out/${TAG}_order${m}_trace_report.txt 2>&1
  head -11 gpurun_out/${TAG}_order${m}_trace_report.txt
done
export FTB200_P2P_ORDER=${MODES%% *}
run > gpurun_out/${TAG}_scale${N}.json 2> gpurun_out/${TAG}_scale${N}.err
line gpurun_out/${TAG}_scale${N}.json validated
