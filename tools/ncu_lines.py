#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts and stall samples from
`ncu -i rep --page source --print-source cuda,sass --csv`.
    python tools/ncu_lines.py both.csv [file-substring] [top-N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_file, out, hdr = "", [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 9 or not r[0].strip().isdigit():
        continue
    try:
        out.append((cur_file, int(r[0]), r[1], float(r[7]), float(r[6])))
    except ValueError:
        pass
tot_i = sum(o[3] for o in out); tot_s = sum(o[4] for o in out)
print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
sel = [o for o in out if want in o[0] or not o[0]]
for f, ln, src, ins, smp in sorted(sel, key=lambda o: -o[3])[:topn]:
    print(f"{ln:5d} {100*ins/tot_i:6.2f}% inst {100*smp/max(tot_s,1):6.2f}% samples  {src.strip()[:110]}")
