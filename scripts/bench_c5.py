"""BASELINE config 5: a batch of small boolean front ends (icosphere k=4 pairs, 5,120 + 5,120
triangles each, seeded offsets) -- "replicas only" (SURVEY 8e): jobs are independent, so they
are spread over host threads, each with its own context (= its own CUDA streams); no collective.

    python scripts/bench_c5.py [--jobs 200] [--threads 4] [--check 8]

Each job goes through the host-buffer C ABI: sb_mesh_create x2, sb_front_end, results to host.
--check N verifies the first N jobs against the CPU oracle.
"""
import argparse, sys, threading, time
import numpy as np
sys.path.insert(0, ".")
import torch
import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

ap = argparse.ArgumentParser()
ap.add_argument("--jobs", type=int, default=200)
ap.add_argument("--threads", type=int, default=4)
ap.add_argument("--check", type=int, default=4)
args = ap.parse_args()

jobs = [meshgen.config_c5_job(j) for j in range(args.jobs)]
results = [None] * args.jobs


def worker(tid):
    ctx = sb.Context(0)
    for j in range(tid, args.jobs, args.threads):
        a, b = jobs[j]
        ma, mb = ctx.mesh(*a), ctx.mesh(*b)
        da = torch.empty(len(a[1]), dtype=torch.uint8, device="cuda")
        db = torch.empty(len(b[1]), dtype=torch.uint8, device="cuda")
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        hab, seg = x.hits()
        results[j] = (x.num_candidates, hab, seg, da.cpu().numpy(), db.cpu().numpy())
        x.close(); ma.close(); mb.close()
    ctx.close()


for warm in (True, False):
    t0 = time.perf_counter()
    ts = [threading.Thread(target=worker, args=(t,)) for t in range(args.threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if not warm:
        print("C5: %d jobs (5,120 + 5,120 tris) on %d host threads: %.1f ms total, %.3f ms/job, %.0f jobs/s"
              % (args.jobs, args.threads, dt * 1e3, dt * 1e3 / args.jobs, args.jobs / dt))

if args.check:
    from oracle import Oracle
    O = Oracle.get()
    for j in range(min(args.check, args.jobs)):
        a, b = jobs[j]
        P, hab, seg, ia, ib = results[j]
        ref = O.candidate_pairs(a, b)
        ret, cop, hit, rseg = O.predicate_pairs(a, b, ref)
        assert P == len(ref) and np.array_equal(hab, ref[hit.astype(bool)]) and seg.tobytes() == rseg[hit.astype(bool)].tobytes()
        oa, _, _ = O.classify(b, O.centroids(*a)); ob, _, _ = O.classify(a, O.centroids(*b))
        assert np.array_equal(ia, oa) and np.array_equal(ib, ob)
    print("checked %d jobs against the oracle: OK" % min(args.check, args.jobs))
