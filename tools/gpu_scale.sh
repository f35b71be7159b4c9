#!/bin/bash
# Multi-GPU measurement session (run under `gpurun --gpus N`): bench lines into gpurun_out/r02_s_<tag>.json
#   tools/gpu_scale.sh <tag> <nproc> [bench.py arguments]
tag=$1; n=$2; shift 2
if [ "$n" = 1 ]; then
  timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r02_s_$tag.json 2> gpurun_out/r02_s_$tag.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $n "$@" > gpurun_out/r02_s_$tag.json 2> gpurun_out/r02_s_$tag.err
fi
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = [json.loads(l) for l in open("gpurun_out/r02_s_%s.json" % tag) if l.startswith("{")][-1]
    print("%-24s N=%d value %.4e  ms/step %.4f  e2e %.3e  valid %s %s" % (tag, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("valid"), json.dumps(d.get("validation"))[:160]))
except Exception as e:
    print(tag, "FAILED", e, open("gpurun_out/r02_s_%s.err" % tag).read()[-400:])
PY
