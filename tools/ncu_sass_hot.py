#!/usr/bin/env python
"""Per-opcode and per-instruction view of an `ncu --page source --csv` export (SASS view).
    python tools/ncu_sass_hot.py src.csv [top-N instructions by stall samples]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_samp = sum(num(r, "# Samples") for r in data)
by_op = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in data:
    src = r[ci["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    by_op[op][0] += num(r, "Instructions Executed"); by_op[op][1] += num(r, "# Samples"); by_op[op][2] += num(r, "L1 Wavefronts Shared")
print(f"total warp instructions {tot_inst:.3e}, samples {tot_samp:.0f}")
print(f"{'opcode':18s} {'inst %':>7s} {'samples %':>9s} {'smem wavefronts':>16s}")
for op, (i, s, w) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{op:18s} {100*i/tot_inst:7.2f} {100*s/max(tot_samp,1):9.2f} {w:16.3e}")
print("\nhottest instructions by stall samples:")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    print(f"{r[ci['Address']][-5:]:6s} {100*num(r,'# Samples')/max(tot_samp,1):6.2f}%  exec {num(r,'Instructions Executed'):.2e}  wf {num(r,'L1 Wavefronts Shared'):.2e}/{num(r,'L1 Wavefronts Shared Ideal'):.2e}  {r[ci['Source']].strip()[:90]}")
