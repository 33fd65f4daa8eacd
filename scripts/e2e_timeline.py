"""Dev tool: host-side and device-side timeline of one host-buffer (e2e) step at C3."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, solidboolean_b200 as sb
from solidboolean_b200 import meshgen
a, b = meshgen.config_c3()
ctx = sb.Context(0); ctx.enable_timing(True)
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (a[0], a[1].view(np.int32), b[0], b[1].view(np.int32))]
nA, nB, nVA, nVB = len(a[1]), len(b[1]), len(a[0]), len(b[0])
da = torch.zeros(nA, dtype=torch.uint8, device="cuda"); db = torch.zeros(nB, dtype=torch.uint8, device="cuda")
oa = torch.zeros(nA, dtype=torch.uint8).pin_memory(); ob = torch.zeros(nB, dtype=torch.uint8).pin_memory()
for it in range(5):
    if it == 4:
        os.environ["SB_DEBUG_SPANS"] = "1"
    ctx.reset_timing(); torch.cuda.synchronize()
    t = [time.perf_counter()]
    xa = sb.Mesh.from_pointers(ctx, pin[0].data_ptr(), nVA, pin[1].data_ptr(), nA, build=False, keep=pin); t.append(time.perf_counter())
    xb = sb.Mesh.from_pointers(ctx, pin[2].data_ptr(), nVB, pin[3].data_ptr(), nB, build=False, keep=pin); t.append(time.perf_counter())
    xa.build(); t.append(time.perf_counter())
    xb.build(); t.append(time.perf_counter())
    x = sb.Isect.front_end(xa, xb, da.data_ptr(), db.data_ptr()); t.append(time.perf_counter())
    x.hits(); oa.copy_(da, non_blocking=True); ob.copy_(db, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    ms, _ = ctx.timing()
    x.close(); xa.close(); xb.close()
print("host ms: uploadA %.3f uploadB %.3f buildA %.3f buildB %.3f front_end %.3f results %.3f total %.3f" % tuple(
    [(t[i + 1] - t[i]) * 1e3 for i in range(6)] + [(t[6] - t[0]) * 1e3]))
print(ms)
