"""Per-op comparison of two bench.py --profile-json files: python tools/profile_diff.py a.json b.json [min_us]"""
import json, sys
a, b = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
print(f"forward {a['forward_ms']:.3f} -> {b['forward_ms']:.3f} ms, step {a['step_ms']:.3f} -> {b['step_ms']:.3f} ms")
for x, y in zip(a["ops"], b["ops"]):
    d = (y["ms"] - x["ms"]) * 1e3
    if abs(d) >= thr:
        print(f"{x['name']:34s} {x['ms']*1e3:8.1f} -> {y['ms']*1e3:8.1f} us ({d:+.1f})")
