#!/usr/bin/env python3
"""List the backward branches (loops) of one SASS function dump and histogram the opcodes of the
biggest loop body.  Usage: sass_loops.py file.sass"""
import re, sys, collections
ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for addr, txt in ins:
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?", txt)
    if "BRA" in txt:
        t = re.search(r"0x([0-9a-f]+)", txt)
        if t:
            tgt = int(t.group(1), 16)
            if tgt <= addr:
                loops.append((addr - tgt, tgt, addr))
loops.sort(reverse=True)
print("instructions:", len(ins))
for size, a, b in loops[:8]:
    n = sum(1 for x, _ in ins if a <= x <= b)
    print(f"loop {a:#x}..{b:#x}: {n} instructions")
if loops:
    size, a, b = loops[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
    h = collections.Counter()
    for x, t in ins:
        if a <= x <= b:
            op = t.split()[0]
            if op.startswith("@"):
                op = t.split()[1]
            h[op.split(".")[0]] += 1
    tot = sum(h.values())
    print("loop body histogram (static):", tot)
    for k, v in h.most_common(45):
        print(f"  {k:10s} {v:5d} {100.0*v/tot:5.1f}%")
