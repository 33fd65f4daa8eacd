"""Half-edge stage on one GPU, beside the reference's addUnintersectedTriangles on
the host (dev / measurement tool): python scripts/halfedge_times.py [c2|c3|c4] [reps] [--no-ref]
Prints one JSON line per repetition and a summary with the algorithmic bytes of the stage."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4": meshgen.config_c4}[cfg]()
ctx = sb.Context(0)
ctx.enable_timing(True)
ma, mb = ctx.mesh(*a), ctx.mesh(*b)
x = ma.intersect(mb)
fa, fb = x.face_flags()
best = None
for it in range(reps):
    ctx.reset_timing()
    t0 = time.perf_counter()
    ua = x.uncut(0, 0, 0)
    ub = x.uncut(1, len(a[0]), ua.num_triangles)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, launches = ctx.timing()
    ctx.reset_timing()
    t2 = time.perf_counter()
    la, ga = ua.components(); lb, gb = ub.components()
    cc_wall = (time.perf_counter() - t2) * 1e3
    cc_ms = ctx.timing()[0]["halfedge"]
    t1 = time.perf_counter()
    ka, oa = ua.half_edges(); kb, ob = ub.half_edges()
    ta = ua.triangles()[1]; tb = ub.triangles()[1]
    down = (time.perf_counter() - t1) * 1e3
    rec = dict(it=it, wall_ms=round(wall, 3), device_ms=round(ms["halfedge"], 4), download_ms=round(down, 3),
               launches=launches, triangles=[ua.num_triangles, ub.num_triangles], ok=[ua.ok, ub.ok],
               components_device_ms=round(cc_ms, 4), components_wall_ms=round(cc_wall, 3), groups=[ga, gb])
    print(json.dumps(rec), flush=True)
    if best is None or rec["device_ms"] < best["device_ms"]:
        best = rec
    ua.close(); ub.close()


def bits_for(n):
    b = 1
    while (1 << b) < n:
        b += 1
    return b


# algorithmic bytes: flags + triples in, face / triple / (key, ordinal) out, one
# read + write of (key, ordinal) per 8-bit radix pass over the 2 * bits(vertices) key bits, then the
# link pass (read sorted pairs, write reference key + owner + adjacency)
total = 0
for m, voff, n in ((a, 0, best["triangles"][0]), (b, len(a[0]), best["triangles"][1])):
    nT = len(m[1])
    passes = (2 * bits_for(voff + len(m[0])) + 7) // 8
    total += nT + 12 * nT + n * (4 + 12) + 3 * n * 12 + passes * 3 * n * 24 + 3 * n * (12 + 16)
summary = dict(config=cfg, device_ms=best["device_ms"], wall_ms=best["wall_ms"], algorithmic_bytes=total,
               gbs=round(total / best["device_ms"] / 1e6, 1), components_device_ms=best["components_device_ms"],
               groups=best["groups"])
if "--no-ref" not in sys.argv:
    from oracle import Ref
    if Ref.available():
        R = Ref.get()
        op = R.op(R.mesh(*a), R.mesh(*b))
        t0 = time.perf_counter()
        (ra, rb), _ = op.uncut(fa, fb)
        summary["reference_harness_ms"] = round((time.perf_counter() - t0) * 1e3, 1)   # incl. the harness' sorted dump
        summary["reference_cpu_ms"] = round(op.uncut_ms(0, 0) + op.uncut_ms(0, 1), 1)  # addUnintersectedTriangles x2 alone, 1 core
        summary["reference_equal"] = bool(np.array_equal(ra["keys"], ka) and np.array_equal(rb["owner"], ob))
        ra_l, ra_g = op.uncut_groups(0, best["triangles"][0])
        rb_l, rb_g = op.uncut_groups(1, best["triangles"][1])
        summary["reference_groups_cpu_ms"] = round(op.uncut_ms(1, 0) + op.uncut_ms(1, 1), 1)   # buildFaceGroups x2 alone
        summary["reference_groups_equal"] = bool(np.array_equal(ra_l, la) and np.array_equal(rb_l, lb) and [ra_g, rb_g] == [ga, gb])
print(json.dumps(summary), flush=True)
