#!/bin/bash
# time library variants (experiments/_lib/lib_*.so) with the headline bench: value, ms/step, element / node kernel ms
for f in experiments/_lib/lib_*.so; do
  FTB200_LIB=$PWD/$f timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu $BENCH_ARGS > /tmp/v.json 2>/tmp/v.err
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/v.json'))
    r = d['roofline']
    print("%-40s value %.4e  ms/step %.4f  elem %.4f  node %.4f  part-loop %s" % (sys.argv[1], d['value'], d['ms_per_step'], r['launch_ms'], r.get('k_node', {}).get('launch_ms', 0), (d.get('n1_partitioned_loop') or {}).get('ms_per_step')))
except Exception as e:
    print(sys.argv[1], "failed", e, open('/tmp/v.err').read()[-300:])
PY
done
