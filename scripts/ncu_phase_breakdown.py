#!/usr/bin/env python
"""Attribute the executed instructions of a kernel to its barrier-separated phases.

  ncu -i X.ncu-rep --page source --csv --print-source sass > X_sass.csv
  python scripts/ncu_phase_breakdown.py X_sass.csv [units]

Splits the SASS listing of the FIRST kernel in the file at every BAR.SYNC and prints, per
segment: warp instructions executed, lane (thread) instructions, stall samples, top opcodes.
`units` (pixels x scales x images of the launch) turns totals into per-unit figures."""
import csv
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    segs, cur = [], dict(inst=0, thr=0, samp=0, ops=Counter(), n=0, first=None)
    for r in rows[hdr_i + 1:end]:
        if len(r) < len(hdr):
            continue
        sass = r[col["Source"]].strip()
        try:
            inst = int(r[col["Instructions Executed"]]); thr = int(r[col["Thread Instructions Executed"]])
            samp = int(r[col["# Samples"]])
        except ValueError:
            continue
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        op = op.split(".")[0]
        cur["inst"] += inst; cur["thr"] += thr; cur["samp"] += samp; cur["ops"][op] += inst; cur["n"] += 1
        if cur["first"] is None:
            cur["first"] = r[col["Address"]]
        if sass.startswith("BAR.SYNC") or " BAR.SYNC" in sass:
            segs.append(cur)
            cur = dict(inst=0, thr=0, samp=0, ops=Counter(), n=0, first=None)
    segs.append(cur)
    tot_i = sum(s["inst"] for s in segs); tot_t = sum(s["thr"] for s in segs); tot_s = sum(s["samp"] for s in segs)
    print(f"total: {tot_i} warp-inst, {tot_t} lane-inst, {tot_s} samples" +
          (f"; per unit: {tot_i / units:.2f} warp-inst, {tot_t / units:.1f} lane-inst" if units else ""))
    for k, s in enumerate(segs):
        if s["inst"] == 0:
            continue
        top = ", ".join(f"{o}:{100 * c / s['inst']:.0f}%" for o, c in s["ops"].most_common(8))
        pu = f" | {s['inst'] / units:6.2f} w-inst/unit {s['thr'] / units:7.1f} lane-inst/unit" if units else ""
        print(f"seg {k:2d} sass={s['n']:5d} inst={100 * s['inst'] / tot_i:5.1f}% samples={100 * s['samp'] / max(tot_s, 1):5.1f}%{pu} | {top}")


if __name__ == "__main__":
    main()
