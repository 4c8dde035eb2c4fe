#!/usr/bin/env python3
"""Stall samples of an .ncu-rep's hottest loop, instruction by instruction in address order (source page):
    ncu_loop.py file.ncu-rep [min-executions-fraction]
Prints every SASS instruction whose execution count is within the given fraction (default 0.5) of the most executed one,
with its share of all stall samples and the dominant stall reasons -- the per-step timeline of a persistent kernel."""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ex = [int(r[col["Instructions Executed"]] or 0) for r in rows[1:]]
mx = max(ex)
tot = sum(int(r[col["Warp Stall Sampling (All Samples)"]] or 0) for r in rows[1:])
print(f"# total samples {tot}, max executions {mx}")
acc = 0
for r, e in zip(rows[1:], ex):
    if e < frac * mx:
        continue
    s = int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    acc += s
    rs = sorted(((int(r[col[k]] or 0), k[6:]) for k in reasons), reverse=True)[:2]
    why = " ".join(f"{k}:{100*v/tot:.1f}" for v, k in rs if v)
    print(f"{100*s/tot:5.1f}% {r[col['Source']].strip()[:70]:70s} {why}")
print(f"# loop share of all samples: {100*acc/tot:.1f}%")

# summary by stall reason over the loop, and per segment between barriers
seg, segs = {}, []
tot_r = {}
for r, e in zip(rows[1:], ex):
    if e < frac * mx:
        continue
    for k in reasons:
        v = int(r[col[k]] or 0)
        seg[k[6:]] = seg.get(k[6:], 0) + v
        tot_r[k[6:]] = tot_r.get(k[6:], 0) + v
    if "BAR.SYNC" in r[col["Source"]] or "BRA" in r[col["Source"]]:
        segs.append(seg)
        seg = {}
print("# by reason (% of all samples): " + " ".join(f"{k}:{100*v/tot:.1f}" for k, v in sorted(tot_r.items(), key=lambda kv: -kv[1]) if v))
for i, sg in enumerate(segs):
    s = sum(sg.values())
    print(f"# segment {i}: {100*s/tot:.1f}%  " + " ".join(f"{k}:{100*v/tot:.1f}" for k, v in sorted(sg.items(), key=lambda kv: -kv[1])[:5] if v))
