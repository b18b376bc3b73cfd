"""Executed warp-instructions by SASS opcode from an ncu source-page CSV (developer tool).

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_opmix.py src.csv [top]
"""
import csv
import sys
from collections import Counter

top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ops, smp = Counter(), Counter()
tot = 0
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if len(r) < 8 or r[0] != "" or not r[2].startswith("0x"):
        continue
    try:
        n, s = int(r[7]), int(r[4])
    except ValueError:
        continue
    toks = r[3].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LD", "ST")) else "")
    ops[op] += n
    smp[op] += s
    tot += n
print(f"total warp instructions {tot}")
ts = sum(smp.values()) or 1
for op, n in ops.most_common(top):
    print(f"{op:22s} {n:12d} {100.0 * n / tot:5.1f}% inst {100.0 * smp[op] / ts:5.1f}% samples")
