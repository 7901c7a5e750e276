#!/usr/bin/env python3
"""Executed warp instructions per CUDA source line: joins `nvdisasm -g -c` of the cubin with the SASS page of an
ncu report (same build!) by instruction address.
usage: ncu_lines.py sass.csv disasm.txt <mangled-substring> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r)
hdr = rows[h]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
cnt = {}
base = None
for r in rows[h + 1:]:
    if len(r) <= it: continue
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None: base = a
        off = a - base
        if off in cnt: break      # the page lists the function twice
        cnt[off] = (int(r[ie] or 0), int(r[it] or 0))
    except ValueError: pass
key = sys.argv[3]
line, on = None, False
per = collections.defaultdict(lambda: [0, 0, 0])
for l in open(sys.argv[2]):
    if l.startswith("//---") : on = key in l
    if not on: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m: line = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and line:
        off = int(m.group(1), 16)
        if off in cnt:
            p = per[line]; p[0] += cnt[off][0]; p[1] += cnt[off][1]; p[2] += 1
tot = sum(v[0] for v in per.values())
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
print("total executed warp instr %.4g" % tot)
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} share={100*v[0]/tot:5.2f}%  sass={v[2]:4d} thr={v[1]/max(v[0],1):5.1f}")
