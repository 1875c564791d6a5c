"""Top SASS instructions (by executed count and by stall samples) of one launch in an ncu source-page CSV."""
import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
launches, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; launches.append(cur); continue
    if cur is not None: cur.append(r)
L = launches[which]; hdr = L[0]; data = L[1:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda d, k: int(float(d[ix[k]] or 0))
tot = sum(I(d, 'Instructions Executed') for d in data); samp = sum(I(d, '# Samples') for d in data)
print(len(launches), 'launches; this one: total warp-inst', tot, 'samples', samp, 'sass lines', len(data))
print("---- top by instructions executed")
for d in sorted(data, key=lambda d: -I(d, 'Instructions Executed'))[:n]:
    print(f"{I(d,'Instructions Executed'):>10d} {100*I(d,'Instructions Executed')/tot:5.1f}% samp {100*I(d,'# Samples')/max(1,samp):5.1f}%  {d[ix['Source']][:110]}")
print("---- top by stall samples")
for d in sorted(data, key=lambda d: -I(d, '# Samples'))[:n//2]:
    print(f"samp {100*I(d,'# Samples')/max(1,samp):5.1f}% inst {100*I(d,'Instructions Executed')/tot:5.1f}%  {d[ix['Source']][:110]}")
