"""Per-CUDA-source-line executed warp instructions and stall samples from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_cs.csv
usage: python profiles/ncu_lines.py X_cs.csv [min_instr] [min_stall_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iex, ist = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
min_ex = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
min_st = int(sys.argv[3]) if len(sys.argv) > 3 else 80
data = []
for r in rows[3:]:
    if len(r) > iex and r[0].isdigit() and r[2] == '-':  # the per-line aggregate rows
        try:
            data.append((int(r[0]), int(r[iex]), int(r[ist]), r[1].strip()))
        except ValueError:
            pass
tot, tots = sum(d[1] for d in data), max(sum(d[2] for d in data), 1)
print("total warp instructions", tot, "stall samples", tots)
for ln, ex, st, src in sorted(data):
    if ex >= min_ex or st >= min_st:
        print(f"{ln:5d} {ex:9d} {100 * ex / tot:5.1f}%  stall {100 * st / tots:5.1f}%  {src[:105]}")
