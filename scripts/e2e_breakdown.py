"""Wall-clock breakdown of the host-buffer (e2e) path: python scripts/e2e_breakdown.py [c3]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import solidboolean_b200 as sb
from solidboolean_b200 import meshgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4": meshgen.config_c4}[cfg]()
ctx = sb.Context(0)
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (a[0], a[1].view(np.int32), b[0], b[1].view(np.int32))]
nVA, nA, nVB, nB = len(a[0]), len(a[1]), len(b[0]), len(b[1])
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    t = [T()]
    xa = sb.Mesh.from_pointers(ctx, pin[0].data_ptr(), nVA, pin[1].data_ptr(), nA, build=True, keep=pin); t.append(T())
    xb = sb.Mesh.from_pointers(ctx, pin[2].data_ptr(), nVB, pin[3].data_ptr(), nB, build=True, keep=pin); t.append(T())
    x = xa.intersect(xb); t.append(T())
    hab, hseg = x.hits(); t.append(T())
    ia, _ = xa.classify_faces_against(xb); t.append(T())
    ib, _ = xb.classify_faces_against(xa); t.append(T())
    x.close(); xa.close(); xb.close(); t.append(T())
    names = ["createA", "createB", "intersect", "hits", "classA", "classB", "destroy"]
    print(it, " ".join("%s=%.3f" % (n, (t[i + 1] - t[i]) * 1e3) for i, n in enumerate(names)), "total=%.3f" % ((t[-1] - t[0]) * 1e3), flush=True)
# raw H2D speed for reference
d = torch.empty(pin[0].numel(), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(pin[0].view(-1), non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("H2D pinned %.1f MB in %.3f ms = %.1f GB/s" % (pin[0].numel() * 8 / 1e6, dt * 1e3, pin[0].numel() * 8 / dt / 1e9))
