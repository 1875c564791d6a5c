"""Per-CUDA-line summary of an ncu report: tools/ncu_lines.py <report.ncu-rep> [top N]
(samples, instructions, dominant stall reasons of the source lines that collect the most warp-stall samples)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]; ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    def g(k):
        try:
            return float(r[ix[k]] or 0)
        except ValueError:
            return 0.0
    st = sorted(((g(s), s[6:]) for s in stalls), reverse=True)[:3]
    lines.append((g("# Samples"), g("Instructions Executed"), int(r[0]), r[1].strip()[:110], st))
tot_s = sum(l[0] for l in lines); tot_i = sum(l[1] for l in lines)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
for s, n, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{s / tot_s * 100:5.1f}% smp {n / tot_i * 100:5.1f}% ins  L{ln:4d} {src}\n        " + ", ".join(f"{k} {v:.0f}" for v, k in st if v))
