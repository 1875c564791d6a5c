import json, sys
from collections import defaultdict
d=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/profile_n256.json'))
print('forward_ms', round(d['forward_ms'],3), 'step_ms', round(d['step_ms'],3))
ops=d['ops']
kinds={0:'stem',1:'conv',2:'dw',3:'pool',4:'attn',5:'decode'}
agg=defaultdict(float)
for o in ops: agg[kinds[o['kind']]]+=o['ms']
print({k:round(v,3) for k,v in agg.items()})
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
for o in sorted(ops,key=lambda o:-o['ms'])[:n]:
    bw = o['mbytes']/o['ms'] if o['ms']>0 else 0
    tf = o['gflop']/o['ms'] if o['ms']>0 else 0
    print(f"{o['name']:36s} {kinds[o['kind']]:6s} {o['ms']:.3f} ms  {o['mbytes']:8.1f} MB {bw:7.0f} GB/s  {o['gflop']:7.1f} GF {tf:6.1f} TF/s")
