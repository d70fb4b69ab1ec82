#!/usr/bin/env python
"""Per-source-line executed instructions of a kernel from
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
Usage: python scripts/ncu_line_breakdown.py X_src.csv [units] [min_pct]"""
import csv
import sys


def main():
    path = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    rows = list(csv.reader(open(path)))
    out, fpath, func, col, seen_func = [], None, None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            func = r[1]
            if seen_func is None:
                seen_func = func
            continue
        if r[0] == "Line No":
            col = {h: i for i, h in enumerate(r)}; continue
        if col is None or func != seen_func or not r[0].strip().isdigit():
            continue
        try:
            inst = int(r[col["Instructions Executed"]]); thr = int(r[col["Thread Instructions Executed"]])
            samp = int(r[col["# Samples"]])
        except (ValueError, KeyError):
            continue
        if inst:
            out.append((fpath, int(r[0]), r[1].strip(), inst, thr, samp))
    tot = sum(o[3] for o in out); tots = sum(o[5] for o in out)
    print(f"kernel: {seen_func}\ntotal {tot} warp-inst" + (f" = {tot / units:.2f}/unit" if units else ""))
    for f, ln, src, inst, thr, samp in out:
        if 100.0 * inst / tot >= min_pct:
            pu = f"{inst / units:5.2f}" if units else ""
            print(f"{f:16s}:{ln:4d} {100 * inst / tot:5.1f}% {pu} smp {100 * samp / max(tots, 1):4.1f}% | {src[:110]}")


if __name__ == "__main__":
    main()
