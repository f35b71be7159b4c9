import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print("value=%.3e ms/step=%.4f elem_ms=%.4f node_ms=%.4f fp64frac=%.3f e2e=%.3e"%(d['value'],d['ms_per_step'],r['launch_ms'],r.get('k_node',{}).get('launch_ms',0.0),r['frac'],d['e2e']['value']))
