"""Dev tool: host->device copy rates from pinned memory on this box (sizes of the C3 upload), per CPU affinity."""
import os, sys, time, subprocess
import torch
def run(tag):
    sizes = [15728688, 15728640, 12582912, 12582912]
    host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
    for h in host: h.fill_(1)
    dev = [torch.empty(s, dtype=torch.uint8, device="cuda") for s in sizes]
    big_h = torch.empty(sum(sizes), dtype=torch.uint8).pin_memory(); big_h.fill_(1)
    big_d = torch.empty(sum(sizes), dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    def one():
        big_d.copy_(big_h, non_blocking=True)
    def four():
        for h, d in zip(host, dev): d.copy_(h, non_blocking=True)
    def two_streams():
        with torch.cuda.stream(s1):
            dev[0].copy_(host[0], non_blocking=True); dev[1].copy_(host[1], non_blocking=True)
        with torch.cuda.stream(s2):
            dev[2].copy_(host[2], non_blocking=True); dev[3].copy_(host[3], non_blocking=True)
    tot = sum(sizes) / 1e6
    for name, fn in (("one 56.6 MB copy", one), ("four copies, one stream", four), ("four copies, two streams", two_streams)):
        ms = timed(fn)
        print("%-22s %-28s %.3f ms  %.1f GB/s" % (tag, name, ms, tot / ms), flush=True)
    # device -> host, 2.9 MB
    dh = torch.empty(2897296, dtype=torch.uint8).pin_memory(); dd = torch.empty(2897296, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: dh.copy_(dd, non_blocking=True))
    print("%-22s %-28s %.3f ms  %.1f GB/s" % (tag, "D2H 2.9 MB", ms, 2.897 / ms), flush=True)
if len(sys.argv) > 1:
    cpus = [int(x) for x in sys.argv[1].split(",")]
    os.sched_setaffinity(0, cpus)
    run("cpus %s.." % sys.argv[1][:12])
else:
    print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i 'numa\\|model name\\|^CPU(s)'; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c", shell=True, capture_output=True, text=True).stdout)
    print("affinity now:", sorted(os.sched_getaffinity(0)))
    run("default")
