#!/usr/bin/env python3
"""Instruction mix of the innermost hot loop (the backward branch whose body holds the most LDS.128) of a kernel.
usage: sass_loop.py <object> <mangled-substring>"""
import re, subprocess, sys, collections
obj, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = key in line
        continue
    if not on: continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s.*0x([0-9a-f]+)", t)
    if t.startswith("@") or "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if "BRA" in t and m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                n_lds = sum("LDS.128" in x for _, x in body)
                if n_lds and (best is None or len(body) < len(best[0]) or False):
                    if best is None or n_lds >= best[1]: best = (body, n_lds)
if not best: sys.exit("no loop found")
body, n_lds = best
c = collections.Counter()
for _, t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    c[t.split()[0].split(".")[0]] += 1
alu = sum(v for k, v in c.items() if k in ("LOP3", "SHF", "IADD3", "LEA", "SEL", "VIADD", "PRMT", "ISETP", "IADD", "VIMNMX", "POPC") )
print(f"loop {len(body)} instr, LDS.128 x{n_lds}")
for k, v in c.most_common(): print(f"  {k:10s} {v}")
print("alu-pipe (LOP3 SHF IADD3 LEA SEL VIADD PRMT ISETP):", alu, " per LDS(column-quarter):", alu / n_lds)
