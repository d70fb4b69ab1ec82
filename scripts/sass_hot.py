#!/usr/bin/env python3
"""Histogram of the hot row loop of a march kernel SASS dump: the largest loop between lo and hi
instructions, with its nested small loops (flush) excluded.  Usage: sass_hot.py f.sass [lo hi rows]"""
import re, sys, collections
ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 3
loops = []
for a, t in ins:
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m:
            tg = int(m.group(1), 16)
            if tg <= a:
                loops.append((tg, a, sum(1 for x, _ in ins if tg <= x <= a)))
cands = [l for l in loops if lo <= l[2] <= hi]
hot = max(cands, key=lambda l: l[2])
inner = [l for l in loops if hot[0] < l[0] and l[1] < hot[1] and l[2] < 400]
h = collections.Counter(); n = 0
for a, t in ins:
    if hot[0] <= a <= hot[1] and not any(x <= a <= y for x, y, _ in inner):
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        h[op.split(".")[0]] += 1; n += 1
print(f"hot loop {hot[0]:#x}..{hot[1]:#x}: {hot[2]} instr, {len(inner)} inner loops excluded -> {n} = {n/rows:.0f}/row")
for k, v in h.most_common(60):
    print(f"  {k:10s} {v:5d} {v/rows:6.1f}/row")
