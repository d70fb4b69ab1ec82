#!/usr/bin/env python3
"""Sum of the control-code stall counts (minimum issue-to-issue cycles of a lone warp) and scoreboard waits of the
row loop of a kernel.  Usage: sass_stallsum.py file.sass"""
import re, sys, collections
lines = open(sys.argv[1]).read().split('\n')
ins = []
i = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            ins.append((int(m.group(1), 16), m.group(2), (w >> 105) & 0xf, (w >> 109) & 1, (w >> 110) & 7, (w >> 113) & 7, (w >> 116) & 0x3f))
            i += 2
            continue
    i += 1
loops = []
for a, t, *_ in ins:
    if 'BRA' in t:
        mm = re.search(r'0x([0-9a-f]+)', t)
        if mm and int(mm.group(1), 16) <= a:
            tg = int(mm.group(1), 16)
            loops.append((tg, a, sum(1 for x in ins if tg <= x[0] <= a)))
hot = min([l for l in loops if l[2] > 150], key=lambda l: l[2])
body = [x for x in ins if hot[0] <= x[0] <= hot[1]]
print("row loop: %d instructions, sum of stall counts %d cycles, %d instructions wait on a scoreboard" %
      (len(body), sum(x[2] for x in body), sum(1 for x in body if x[6])))
h = collections.Counter(x[2] for x in body)
print("stall-count histogram:", dict(sorted(h.items())))
