#!/bin/bash
# automask pre-pass (photometric_min forward): kernel times under `ncu --metrics gpu__time_duration.sum` (serialised, cold
# clocks policy of the profiling recipe) for a list of build variants, e.g. "-DMD2_PM_BULK=0" "-DMD2_PM_MINB=3"
mkdir -p gpurun_out
B="import importlib.util; spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()"
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "$B" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[variant '$V']"
  for SH in "640 192 12 3" "416 128 64 3" "1024 320 4 3"; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:photomin --csv --log-file gpurun_out/pm_ncu.csv python scripts/exp/pm_ncu.py $SH > /dev/null 2>&1
    python - "$SH" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/pm_ncu.csv")) if len(r) > 5 and r[0].isdigit()]
t = sorted(float(r[-1].replace(",", "")) for r in rows)
u = rows[0][-2] if rows else "?"
print("  ", sys.argv[1], ": n", len(t), "median", t[len(t) // 2] if t else None, "min", t[0] if t else None, u)
PY
  done
done
