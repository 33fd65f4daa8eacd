"""Stage timing breakdown on one GPU (dev tool): python scripts/stage_times.py [c2|c3|c4] [reps]"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t0 = time.time()
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4": meshgen.config_c4,
        "c4k8": lambda: meshgen.config_c4(k=8)}[cfg]()   # c4k8: 1,310,720 x2 near-coincident, ~9 M candidate pairs
print("gen %.2fs  A %d tris  B %d tris" % (time.time() - t0, len(a[1]), len(b[1])), flush=True)
import torch
ctx = sb.Context(0)
ctx.enable_timing("--no-timing" not in sys.argv)   # --no-timing: no stage events on the streams (wall time only)
ma = ctx.mesh(*a, build=False)
mb = ctx.mesh(*b, build=False)
da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if "--flush" in sys.argv else None   # as bench.py: L2 evicted between steps
for it in range(reps):
    if flush is not None:
        flush.zero_(); torch.cuda.synchronize()
    ctx.reset_timing()
    t0 = time.perf_counter()
    ma.build(); mb.build()
    if "--serial" in sys.argv:
        x = ma.intersect(mb)
        ma.classify_faces_device(mb, da.data_ptr())
        rA = ctx.classify_stats()
        mb.classify_faces_device(ma, db.data_ptr())
        rB = ctx.classify_stats()
    else:
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        rA = rB = ctx.classify_stats()
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, launches = ctx.timing()
    print(json.dumps(dict(it=it, wall_ms=round(wall, 3), stages={k: round(v, 4) for k, v in ms.items()},
                          launches=launches, P=x.num_candidates, H=x.num_hits, raysA=rA, raysB=rB,
                          insideA=int(da.sum()), insideB=int(db.sum()))), flush=True)
    x.close()
print("gridA", ma.grid_info()); print("gridB", mb.grid_info())
