#!/bin/bash
# Dev tool: build time of A alone / B alone / both at C3, per-kernel list of the last builds (graphs off, ncu), parity tests
python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -3
python scripts/build_times.py c3
SB_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_build.csv python scripts/build_times.py c3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_build.csv") if l.startswith('"')))
h = rows[0]; iK = h.index("Kernel Name"); iM = h.index("Metric Name"); iV = h.index("Metric Value"); iI = h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iI], {"k": r[iK]})[r[iM]] = float(r[iV].replace(",", ""))
L = list(d.values())
for x in L[-22:]:
    t = x["gpu__time_duration.sum"]
    print("%-60s %8.1f us  rd %7.1f MB  wr %7.1f MB" % (x["k"][:60], t / 1e3 if t > 1e3 else t, x.get("dram__bytes_read.sum", 0) / 1e6, x.get("dram__bytes_write.sum", 0) / 1e6))
PY
for c in c3 c2; do python scripts/stage_times.py $c 5 2>&1 | grep '"it": 4' | cut -c1-200; done
