"""Dev tool: one rank's share of a multi-GPU step on ONE GPU (no exchange): wall / device time and the stage spans.
    python scripts/shard_timeline.py [n_ranks] [rank] [reps]"""
import os, sys, time
sys.path.insert(0, ".")
import torch, solidboolean_b200 as sb
from solidboolean_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
r = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
a, b = meshgen.config_c3()
ctx = sb.Context(0); ctx.enable_timing(True)
ma = ctx.mesh(*a, build=False); mb = ctx.mesh(*b, build=False)
buf = torch.zeros(len(a[1]) + len(b[1]), dtype=torch.uint8, device="cuda")
ext = torch.cuda.ExternalStream(ctx.stream)
sh = sb.Shard(ma, mb, r, n)
for it in range(reps):
    if it == reps - 1:
        os.environ["SB_DEBUG_SPANS"] = "1"
    ctx.reset_timing()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(ext):
        e0.record(ext)
        buf.zero_()
        x = sh.front_end(buf.data_ptr(), buf.data_ptr() + len(a[1]))
        t1 = time.perf_counter()
        e1.record(ext)
    e1.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, launches = ctx.timing()
    print("it %d  device %.3f ms  host call %.3f ms  wall %.3f ms  launches %d  %s" % (
        it, e0.elapsed_time(e1), (t1 - t0) * 1e3, wall, launches, {k: round(v, 3) for k, v in ms.items() if v}), flush=True)
    x.close()
print(sh.info())
