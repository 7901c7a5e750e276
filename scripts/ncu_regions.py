#!/usr/bin/env python3
"""Where a kernel's executed instructions go: contiguous SASS regions of equal execution count from
`ncu -i X.ncu-rep --page source --csv --print-source sass [--launch-skip N --launch-count 1] > sass.csv`.
usage: ncu_regions.py sass.csv [min_share_percent]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
h = next(i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r)
hdr = rows[h]
ia, isrc, ie, it = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = []
for r in rows[h + 1:]:
    if len(r) <= it: continue
    try: data.append((r[ia], r[isrc], int(r[ie] or 0), int(r[it] or 0)))
    except ValueError: pass
tot = sum(d[2] for d in data); tt = sum(d[3] for d in data)
print("SASS instructions", len(data), " executed warp-instr %.4g" % tot, " avg active threads %.2f" % (tt / tot))
segs, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][2] - data[i - 1][2]) > 0.02 * max(data[i - 1][2], 1):
        segs.append((start, i - 1, sum(x[2] for x in data[start:i]), sum(x[3] for x in data[start:i])))
        start = i
for a, b, c, t in segs:
    if c > tot * minshare / 100:
        print(f"{a:5d}-{b:5d} n={b-a+1:4d} exec/instr={c//(b-a+1):>12d} share={100*c/tot:5.1f}% thr={t/max(c,1):5.1f}  {data[a][1][:70]}")
