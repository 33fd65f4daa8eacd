#!/bin/bash
# Dev tool: per-kernel build times (ncu launch list, graphs off) of every library variant
for f in solidboolean_b200/lib/libsolidboolean_b200.so solidboolean_b200/lib/variants/libsb_*.so; do echo $f
SB_LIB_PATH=$PWD/$f python scripts/build_times.py c3
SB_LIB_PATH=$PWD/$f SB_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lb.csv python scripts/build_times.py c3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/lb.csv") if l.startswith('"')))
h = rows[0]; iK = h.index("Kernel Name"); iV = h.index("Metric Value")
L = [(r[iK], float(r[iV].replace(",", ""))) for r in rows[1:]]
for k, t in L[-21:-10]:
    print("   %-50s %8.1f us" % (k[:50], t / 1e3 if t > 1e3 else t))
PY
done
