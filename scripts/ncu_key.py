#!/usr/bin/env python3
"""Key figures of one kernel from an ncu report (raw page).  Usage: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, v = rows[0], rows[-1]
d = dict(zip(hdr, v))
def f(k):
    try: return float(d[k])
    except Exception: return float("nan")
print("kernel", d.get("Kernel Name", "")[:60], "grid", d.get("launch__grid_size"), "regs", d.get("launch__registers_per_thread"))
el, act = f("sm__cycles_elapsed.max"), f("sm__cycles_active.avg")
inst = f("smsp__inst_executed.sum")
print(f"duration {f('gpu__time_duration.sum'):.1f} us  elapsed {el:.0f} cyc  sm active avg {act:.0f} ({100*act/el:.0f}%)  min {f('sm__cycles_active.min'):.0f} max {f('sm__cycles_active.max'):.0f}")
print(f"warp-inst {inst/1e6:.2f} M = {inst/592:.0f} per scheduler; issue active {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}% of active; IPC/sched over elapsed {inst/592/el:.3f}")
print(f"warps active {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}%  fma pipe {f('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f}%  lsu {f('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f}%")
print(f"dram read {f('dram__bytes_read.sum')} write {f('dram__bytes_write.sum')}  l1 hit {f('l1tex__t_sector_hit_rate.pct'):.1f}%  l2 hit {f('lts__t_sector_hit_rate.pct'):.1f}%")
