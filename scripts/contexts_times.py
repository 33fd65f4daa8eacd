"""Per-triangle intersection contexts on one GPU beside the oracle port on one host
core (dev / measurement tool): python scripts/contexts_times.py [c2|c3|c4|c4k8] [reps]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4": meshgen.config_c4,
        "c4k8": lambda: meshgen.config_c4(k=8)}[cfg]()
ctx = sb.Context(0)
ctx.enable_timing(True)
ma, mb = ctx.mesh(*a), ctx.mesh(*b)
x = ma.intersect(mb)
lib = x.lib
best = None
for it in range(reps):
    ctx.reset_timing()
    t0 = time.perf_counter()
    hs = []
    for w in (0, 1):
        h = sb._vp()
        sb._check(lib.sb_isect_contexts(x.h, w, sb.C.byref(h)))
        hs.append(h)
    wall = (time.perf_counter() - t0) * 1e3
    ms, launches = ctx.timing()
    counts = []
    for h in hs:
        nc, npt, ne = sb._sz(0), sb._sz(0), sb._sz(0)
        lib.sb_cuts_counts(h, sb.C.byref(nc), sb.C.byref(npt), sb.C.byref(ne))
        counts.append([nc.value, npt.value, ne.value])
        lib.sb_cuts_destroy(h)
    rec = dict(it=it, device_ms=round(ms["contexts"], 4), wall_ms=round(wall, 3), launches=launches, hits=x.num_hits,
               contexts_points_relations=counts)
    print(json.dumps(rec), flush=True)
    if best is None or rec["device_ms"] < best["device_ms"]:
        best = rec
# algorithmic bytes per side: hits in (8 H) + segments in (48 H) + contexts out (4 C) + CSR (8 C) + points (24 P) + relations (8 E)
H = x.num_hits
total = sum(56 * H + 12 * c + 24 * p + 8 * e for c, p, e in best["contexts_points_relations"])
summary = dict(config=cfg, hits=H, device_ms=best["device_ms"], wall_ms=best["wall_ms"], algorithmic_bytes=total,
               gbs=round(total / best["device_ms"] / 1e6, 2))
from oracle import Oracle
O = Oracle.get()
hab, seg = x.hits()
t0 = time.perf_counter()
ref = [O.cut_contexts(hab, seg, w) for w in (0, 1)]
summary["oracle_port_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
dev = [x.contexts(w) for w in (0, 1)]
summary["identical"] = bool(all(np.array_equal(d[k], r[k]) for d, r in zip(dev, ref) for k in ("tri", "point_start", "edge_start", "edges"))
                            and all(d["points"].tobytes() == r["points"].tobytes() for d, r in zip(dev, ref)))
print(json.dumps(summary), flush=True)
