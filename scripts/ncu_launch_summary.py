#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
usage: ncu_launch_summary.py X.csv > summary.md"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum": continue
    k = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("trpa::", "")
    tot[k] += float(r[iv].replace(",", "")) / 1e6
    cnt[k] += 1
allms = sum(tot.values())
print("kernel | launches | total_ms | share")
print("---|---|---|---")
for k, v in tot.most_common():
    print(f"{k} | {cnt[k]} | {v:.3f} | {100 * v / allms:.1f}%")
print(f"all | {sum(cnt.values())} | {allms:.3f} | 100%")
