"""Dev tool: the host-buffer (e2e) step of bench.py at C3, on its own: python scripts/e2e_quick.py [steps]
(sb_mesh_update x2 from pinned memory + sb_mesh_build x2 + sb_front_end + results to pinned memory; L2 flushed between steps)"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, solidboolean_b200 as sb
from solidboolean_b200 import meshgen
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
a, b = meshgen.config_c3()
ctx = sb.Context(0)
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (a[0], a[1].view(np.int32), b[0], b[1].view(np.int32))]
nA, nB = len(a[1]), len(b[1])
ma = sb.Mesh.from_pointers(ctx, pin[0].data_ptr(), len(a[0]), pin[1].data_ptr(), nA, build=False, keep=pin)
mb = sb.Mesh.from_pointers(ctx, pin[2].data_ptr(), len(b[0]), pin[3].data_ptr(), nB, build=False, keep=pin)
da = torch.zeros(nA, dtype=torch.uint8, device="cuda"); db = torch.zeros(nB, dtype=torch.uint8, device="cuda")
oa = torch.zeros(nA, dtype=torch.uint8).pin_memory(); ob = torch.zeros(nB, dtype=torch.uint8).pin_memory()
hab = torch.zeros(2 * 20000, dtype=torch.int32).pin_memory(); hseg = torch.zeros(6 * 20000, dtype=torch.float64).pin_memory()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ts = []
import contextlib
ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
scope = (lambda: torch.cuda.stream(ext)) if os.environ.get("E2E_EXT") else contextlib.nullcontext
if os.environ.get("E2E_SAMPLER"):
    import bench
    smp = bench.ClockSampler(0); smp.start()
for it in range(steps + 3):
    flush.fill_(it & 0xff); torch.cuda.synchronize()
    spans = os.environ.get("E2E_SPANS") and it == steps + 2
    if os.environ.get("E2E_SPANS") == "all" and not spans:   # timing on in every step (the last one prints)
        ctx.enable_timing(True); ctx.reset_timing()
    if spans:   # last step: the stage spans (ms since the reset) on stderr, host time stamps of the calls below
        ctx.enable_timing(True); ctx.reset_timing(); os.environ["SB_DEBUG_SPANS"] = "1"
    t0 = time.perf_counter()
    cm = scope(); cm.__enter__()
    if os.environ.get("E2E_ORDER", "ab") == "ab":
        ma.update(pin[0].data_ptr(), pin[1].data_ptr()); mb.update(pin[2].data_ptr(), pin[3].data_ptr())
        ma.build(); mb.build()
    else:   # the smaller mesh first: its build is over sooner, the larger one's faces are classified as they arrive
        mb.update(pin[2].data_ptr(), pin[3].data_ptr()); ma.update(pin[0].data_ptr(), pin[1].data_ptr())
        mb.build(); ma.build()
    tb = time.perf_counter()
    if os.environ.get("E2E_HOST_API", "1") == "1":
        x = sb.Isect.front_end_host(ma, mb, oa.data_ptr(), ob.data_ptr(), hab.data_ptr(), hseg.data_ptr(), 20000)
        tf = time.perf_counter()
    else:
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        tf = time.perf_counter()
        oa.copy_(da, non_blocking=True); ob.copy_(db, non_blocking=True)
        sb._check(x.lib.sb_isect_hits(x.h, hab.data_ptr(), hseg.data_ptr()))
        torch.cuda.synchronize()
    if os.environ.get("E2E_CLOSE_INSIDE"):
        x.close(); x = None
    cm.__exit__(None, None, None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if spans:
        print("host ms: update+build calls returned %.3f, front_end returned %.3f, results on the host %.3f" % ((tb - t0) * 1e3, (tf - t0) * 1e3, (t1 - t0) * 1e3))
        ctx.timing()
    if x is not None:
        x.close()
    if it >= 3:
        ts.append((t1 - t0) * 1e3)
ts = np.array(ts)
print("e2e ms/step: mean %.3f  median %.3f  min %.3f   inside %d %d  [SB_STREAM_CLASSIFY=%s SB_STREAM_CTAS=%s order %s host api %s close-inside %s ext %s sampler %s]" % (
    ts.mean(), np.median(ts), ts.min(), int(oa.sum()), int(ob.sum()), os.environ.get("SB_STREAM_CLASSIFY", "-"), os.environ.get("SB_STREAM_CTAS", "-"), os.environ.get("E2E_ORDER", "ab"), os.environ.get("E2E_HOST_API", "1"), os.environ.get("E2E_CLOSE_INSIDE", "-"),
    os.environ.get("E2E_EXT", "-"), os.environ.get("E2E_SAMPLER", "-")))
