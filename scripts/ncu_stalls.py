#!/usr/bin/env python3
"""Stall summary of one kernel from an ncu report: totals per stall reason and the top stall sites (SASS).
Usage: ncu_stalls.py report.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(g(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:80])
print("samples", int(tot), "warp-instructions", int(sum(g(r, "Instructions Executed") for r in data)))
print("  ".join(f"{k[6:]}={100 * sum(g(r, k) for r in data) / tot:.1f}%" for k in sorted(keys, key=lambda k: -sum(g(r, k) for r in data)) if sum(g(r, k) for r in data) > 0))
for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:ntop]:
    st = sorted(((k, g(r, k)) for k in keys), key=lambda kv: -kv[1])[:3]
    print(r[ix["Address"]][-5:], f'{int(g(r, "# Samples")):4d}', r[ix["Source"]][:72].ljust(72), " ".join(f"{k[6:]}={int(v)}" for k, v in st if v > 0))
