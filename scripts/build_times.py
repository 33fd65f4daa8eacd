"""Dev tool: build time of mesh A alone, mesh B alone and both together (CUDA events around the calls, after warm-up)
    python scripts/build_times.py [c3|c2|c4k8]"""
import sys
sys.path.insert(0, ".")
import torch, solidboolean_b200 as sb
from solidboolean_b200 import meshgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4k8": lambda: meshgen.config_c4(k=8)}[cfg]()
ctx = sb.Context(0)
ma = ctx.mesh(*a, build=False); mb = ctx.mesh(*b, build=False)
da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda"); db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
ma.build(); mb.build()
x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr()); x.close()   # B becomes a traversal target (LBVH in its build)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=10):
    best = []
    for _ in range(reps):
        flush.fill_(1)
        ctx.synchronize(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ctx.enable_timing(True); ctx.reset_timing()
        fn()
        ctx.synchronize()
        ms, _ = ctx.timing()
        best.append(ms["build"])
    best.sort()
    return best[len(best) // 2]
print("A alone  %.4f ms" % timed(lambda: ma.build()))
print("B alone  %.4f ms" % timed(lambda: mb.build()))
print("A and B  %.4f ms" % timed(lambda: (ma.build(), mb.build())))
