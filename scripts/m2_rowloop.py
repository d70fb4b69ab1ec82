#!/usr/bin/env python3
"""Static opcode histogram of the row loop (the innermost loop with > 150 instructions) of the single-warp marching kernel,
and instructions per source line inside it (nvdisasm -g listing).  Usage: m2_rowloop.py file.sass file.dis"""
import re, sys, collections
ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, t in ins:
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            tg = int(m.group(1), 16)
            loops.append((tg, a, sum(1 for x, _ in ins if tg <= x <= a)))
cands = [l for l in loops if l[2] > 150]
hot = min(cands, key=lambda l: l[2])
h = collections.Counter()
for a, t in ins:
    if hot[0] <= a <= hot[1]:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        h[op.split(".")[0]] += 1
n = sum(h.values())
print(f"row loop {hot[0]:#x}..{hot[1]:#x}: {n} instructions")
print("  " + "  ".join(f"{k} {v}" for k, v in h.most_common(60)))
if len(sys.argv) > 2:
    # per source line
    cur = None; per = collections.Counter(); addr_re = re.compile(r"/\*([0-9a-f]{4,})\*/")
    for line in open(sys.argv[2]):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = addr_re.search(line)
        if m and cur:
            a = int(m.group(1), 16)
            if hot[0] <= a <= hot[1]:
                per[cur] += 1
    src = {}
    for (f, l), v in sorted(per.items()):
        print(f"  {f}:{l}: {v}")
