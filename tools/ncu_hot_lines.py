"""Per-source-line stall samples / instruction counts from an ncu report captured with -lineinfo and
--import-source on (developer tool).

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_hot_lines.py src.csv [top]
"""
import csv
import sys

top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = []
tot_s = tot_i = 0
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No") or r[0] == "" or len(r) < 8:
        continue
    try:
        ln, samples, inst = int(r[0]), int(r[4]), int(r[7])
    except ValueError:
        continue
    lines.append((samples, inst, cur_file, ln, r[1].strip()[:110]))
    tot_s += samples
    tot_i += inst
lines.sort(reverse=True)
print(f"total samples {tot_s}  warp instructions {tot_i}")
for s, i, f, ln, src in lines[:top]:
    print(f"{100.0 * s / max(tot_s, 1):5.1f}% smp {100.0 * i / max(tot_i, 1):5.1f}% inst  {f}:{ln:<4d} {src}")
