#!/usr/bin/env python
"""Dynamic (executed) opcode histogram + per-region totals of the first kernel in
  ncu -i X.ncu-rep --page source --csv > X_sass.csv
Usage: ncu_sass_dyn.py X_sass.csv units [dump]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
ops, tot, thr_tot, ins = Counter(), 0, 0, []
for r in rows[hi + 1:end]:
    if len(r) < len(rows[hi]):
        continue
    s = r[col["Source"]].strip()
    try:
        n = int(r[col["Instructions Executed"]]); t = int(r[col["Thread Instructions Executed"]]); sm = int(r[col["# Samples"]])
    except ValueError:
        continue
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    ops[op.split(".")[0]] += n
    tot += n; thr_tot += t
    ins.append((s, n, t, sm))
print(f"total warp-inst {tot} = {tot/units:.2f}/unit, lane-inst {thr_tot/units:.1f}/unit, avg active lanes {thr_tot/max(tot,1):.1f}")
for k, v in ops.most_common(40):
    print(f"  {k:10s} {v/units:7.3f}/unit {100*v/tot:5.1f}%")
if len(sys.argv) > 3:
    for i, (s, n, t, sm) in enumerate(ins):
        print(f"{i:5d} {n:9d} {t/max(n,1):5.1f} {sm:5d}  {s}")
