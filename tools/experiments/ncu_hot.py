#!/usr/bin/env python3
"""Top stall-sampled SASS instructions of an .ncu-rep (source page):  ncu_hot.py file.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ia, isrc, iall, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(r[iall] or 0), idx, r[ia], r[isrc], int(r[iex] or 0)) for idx, r in enumerate(rows[1:]) if len(r) > iall]
tot = sum(d[0] for d in data)
print(f"# {lines[0][:120]}  total samples {tot}")
for s, idx, addr, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{100*s/tot:6.2f}%  #{idx:5d} ex={ex:9d}  {src.strip()[:110]}")
