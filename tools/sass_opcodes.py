#!/usr/bin/env python3
"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass): which kernels hold tcgen05 (UTCHMMA / UTCBAR),
TMEM traffic (LDTM / STTM), legacy tensor-core HMMA, MUFU, cp.async (LDGSTS), bulk copies (UBLKCP / UTMALDG).

    python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neural-tape-modeling_b200", "libntm_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "HMMA", "MUFU.EX2", "MUFU.RCP", "MUFU.TANH", "MUFU", "FFMA2", "FFMA", "FMUL2",
         "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "SHFL", "DFMA", "F2FP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "MUFU" and op.startswith("MUFU")):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (tools/sass_opcodes.py; static instruction counts, sm_100a)")
    print("# columns: " + " ".join(WATCH) + " | total")
    groups = collections.OrderedDict()
    for (name, c), dn in zip(kernels.items(), demangle):
        base = re.sub(r"<.*", "", dn).replace("void ", "")
        groups.setdefault(base, []).append((dn, c))
    for base, inst in groups.items():
        print(f"\n== {base}: {len(inst)} instance(s)")
        for dn, c in inst:
            args = re.search(r"<(.*)>", dn)
            row = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
            print(f"  <{args.group(1) if args else ''}>  {row} | total={c['_total']}")
    print("\n# summary: kernels holding each Blackwell-specific opcode")
    for w in ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "HMMA", "LDGSTS", "FFMA2"):
        names = sorted({b for b, inst in groups.items() if any(c[w] for _, c in inst)})
        print(f"  {w}: {', '.join(names) if names else '(none)'}")


if __name__ == "__main__":
    main()
