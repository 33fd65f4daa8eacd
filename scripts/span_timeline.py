import sys, time
sys.path.insert(0, ".")
import torch, solidboolean_b200 as sb
from solidboolean_b200 import meshgen
a, b = meshgen.config_c3()
ctx = sb.Context(0); ctx.enable_timing(True)
ma = ctx.mesh(*a, build=False); mb = ctx.mesh(*b, build=False)
da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda"); db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
for it in range(5):
    if it == 4:
        import os; os.environ["SB_DEBUG_SPANS"] = "1"
    ctx.reset_timing()
    t0 = time.perf_counter()
    ma.build(); mb.build()
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, _ = ctx.timing()
    x.close()
print("wall", wall, ms)
