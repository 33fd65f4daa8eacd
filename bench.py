#!/usr/bin/env python
"""Benchmark of the solidboolean intersection front end on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

One "step" = one pass of the hot path over the workload:
build(A) + build(B) + broad phase + predicate + hit compaction + inside/outside
classification of every face of A against B and of B against A.

* value  : intersecting tri-pairs/s with the geometry already resident in HBM
           (per-step CUDA-event time on the library's stream, L2 flushed between
           steps, max over ranks).
* e2e    : the same metric through the host-buffer C ABI (sb_mesh_create from
           pinned host memory -> H2D inside, results read back to the host).
* roofline: the dominant kernel group (classification) against the measured HBM
           copy bandwidth, using the algorithmic-bytes formula of SURVEY 8d.
* cpu_baseline: the reference's own CPU implementation (oracle/_ref, or the C
           port when it is absent) on this box's host cores, bounded sample.

`--impl reference` times only the CPU reference on the same config and metric.
N > 1 (torchrun): mesh A's Morton range and both classification query sets are
sharded over the ranks, mesh B's (and A's) LBVH replicated; hit lists are
all-gathered and the face flags summed over NCCL.  Total work is fixed -> strong
scaling.
"""
from __future__ import annotations

import argparse
import json
import contextlib
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "intersecting_tri_pairs_per_s"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c4k8", "c5"])
    ap.add_argument("--jobs", type=int, default=1000, help="c5: number of 5K+5K-triangle jobs in the batch")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(name):
    from solidboolean_b200 import meshgen
    a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3, "c4": meshgen.config_c4,
            "c4k8": lambda: meshgen.config_c4(k=8)}[name]()
    desc = {
        "c2": "two offset icospheres k=6, 81,920 + 81,920 triangles (BASELINE configs[1])",
        "c3": "icosphere k=8 (1,310,720 tris) vs torus 1024x512 (1,048,576 tris), BASELINE configs[2] = the 1M+1M config the metric is quoted on",
        "c4": "near-coincident icospheres k=7, 327,680 x2 triangles, ~2.3M candidate pairs (BASELINE configs[3] proxy)",
        "c4k8": "near-coincident icospheres k=8, 1,310,720 x2 triangles, ~8.2M candidate pairs (BASELINE configs[3] at its stated size)",
    }[name]
    return a, b, desc


# ----------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        """Clocks / throttle reasons of the samples taken in [t0, t1] (monotonic seconds);
        if the window is shorter than the sampling period, the samples nearest to it."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for (t, r) in self.rows if t0 <= t <= t1]
        note = "samples inside the timed region"
        if not rows and self.rows:
            mid = 0.5 * (t0 + t1)
            rows = [r for (_, r) in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            note = "timed region shorter than the sampling period: nearest samples (GPU under the same load)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "note": note}


# ----------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference; else the C port)

def cpu_reference_step(a, b, sample_points, threads):
    """One pass of the reference front end on the host: prepare() of both meshes, the broad
    phase, the predicate loop and isPointInMesh on the face centroids (sample_points = None:
    ALL of them, nothing extrapolated; a number: that many, spread evenly, and the time scaled).
    Returns (seconds, H, detail dict); detail["measured_s"] is what the clock saw."""
    from oracle import Oracle, Ref
    nA, nB = len(a[1]), len(b[1])
    Q = nA + nB
    full = sample_points is None or sample_points >= Q
    rng = np.random.default_rng(0)

    def pick(n, share):
        return np.arange(n) if full else rng.choice(n, size=min(n, max(1, share)), replace=False)
    if Ref.available():
        R = Ref.get()
        out = [None, None]

        def prep(i, m):
            out[i] = R.mesh(*m)
        t0 = time.perf_counter()
        ts = [threading.Thread(target=prep, args=(i, m)) for i, m in enumerate((a, b))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        t_prepare = time.perf_counter() - t0          # two meshes on two cores (ctypes drops the GIL)
        ma, mb = out
        op = R.op(ma, mb)
        t0 = time.perf_counter()
        pairs = op.search()
        t_search = time.perf_counter() - t0
        ret, cop, hit, seg, ms_pred = op.predicate(pairs)
        t_pred = ms_pred / 1e3
        H = int(hit.sum())
        # isPointInMesh on the face centroids (3 axes each), spread over the host threads
        ca, cb = ma.centroids(), mb.centroids()
        sa = pick(nA, (sample_points or 0) * nA // Q)
        sb_ = pick(nB, (sample_points or 0) * nB // Q)
        jobs = [(1, chunk) for chunk in np.array_split(ca[sa], threads)] + \
               [(0, chunk) for chunk in np.array_split(cb[sb_], threads)]
        inside = [0, 0]
        parts = [[None] * threads, [None] * threads]   # per-face flags, chunk by chunk (full runs: compared with the GPU's)

        def work(k, j):
            flags = op.classify(j[0], j[1])[0]
            parts[1 - j[0]][k % threads] = flags
        t0 = time.perf_counter()
        ws = [threading.Thread(target=work, args=(k, j)) for k, j in enumerate(jobs) if len(j[1])]
        for i in range(0, len(ws), threads):
            [w.start() for w in ws[i:i + threads]]
            [w.join() for w in ws[i:i + threads]]
        t_cls_sample = time.perf_counter() - t0
        n_sample = len(sa) + len(sb_)
        kind = "reference"
        flags_ab = [np.concatenate([x for x in parts[k] if x is not None]) if any(x is not None for x in parts[k])
                    else np.zeros(0, np.uint8) for k in (0, 1)]
        inside = [int(flags_ab[0].sum()), int(flags_ab[1].sum())]
        op.close(); ma.close(); mb.close()
    else:
        O = Oracle.get()
        t0 = time.perf_counter()
        O.tri_boxes(*a); O.tri_boxes(*b); O.normals(*a); O.normals(*b)
        t_prepare = time.perf_counter() - t0
        t0 = time.perf_counter()
        pairs = O.candidate_pairs(a, b)
        t_search = time.perf_counter() - t0
        t0 = time.perf_counter()
        ret, cop, hit, seg = O.predicate_pairs(a, b, pairs)
        t_pred = time.perf_counter() - t0
        H = int(hit.sum())
        ca, cb = O.centroids(*a), O.centroids(*b)
        sa = pick(nA, (sample_points or 0) * nA // Q)
        sb_ = pick(nB, (sample_points or 0) * nB // Q)
        jobs = [(b, 0, chunk) for chunk in np.array_split(ca[sa], threads)] + \
               [(a, 1, chunk) for chunk in np.array_split(cb[sb_], threads)]
        parts = [[None] * threads, [None] * threads]

        def work(k, j):
            parts[j[1]][k % threads] = O.classify(j[0], j[2])[0]
        t0 = time.perf_counter()
        ws = [threading.Thread(target=work, args=(k, j)) for k, j in enumerate(jobs) if len(j[2])]
        for i in range(0, len(ws), threads):
            [w.start() for w in ws[i:i + threads]]
            [w.join() for w in ws[i:i + threads]]
        t_cls_sample = time.perf_counter() - t0
        n_sample = len(sa) + len(sb_)
        kind = "port"
        flags_ab = [np.concatenate([x for x in parts[k] if x is not None]) if any(x is not None for x in parts[k])
                    else np.zeros(0, np.uint8) for k in (0, 1)]
        inside = [int(flags_ab[0].sum()), int(flags_ab[1].sum())]
    t_cls = t_cls_sample * (Q / max(n_sample, 1))
    measured = t_prepare + t_search + t_pred + t_cls_sample
    total = t_prepare + t_search + t_pred + t_cls
    detail = dict(kind=kind, prepare_s=t_prepare, search_s=t_search, predicate_s=t_pred,
                  classify_measured_s=t_cls_sample, classify_points=n_sample, classify_points_total=Q,
                  extrapolated_s=total - measured, measured_s=measured, H=H, P=int(len(pairs)),
                  inside_a=inside[0] if full else None, inside_b=inside[1] if full else None)
    detail["_flags"] = flags_ab if full else None        # numpy arrays, for the caller's comparison (never serialised)
    detail["_hits"] = (pairs[hit.astype(bool)], seg[hit.astype(bool)])
    return total, H, detail


def other_configs(sb, ctx, ext, l2_flush, torch, steps=10, warmup=3):
    """Resident step time (build x2 + front end, CUDA events on the library's stream, L2 flushed between
    steps) of BASELINE configs[1] (C2) and configs[3] at its stated size (C4: 8.2 M candidate pairs), and for C4
    the comparison of the counts with what the unmodified reference left in tests/golden/fullsize.json (the hashes of every output are compared by tests/test_gpu_parity.py)."""
    out = {}
    pins = {}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize.json")) as f:
            pins = json.load(f)
    except Exception:
        pass

    for name in ("c2", "c4k8"):
        a, b, desc = workload(name)
        ma, mb = ctx.mesh(*a, build=False), ctx.mesh(*b, build=False)
        fa = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
        fb = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
        res = [0, 0]

        def step():
            with torch.cuda.stream(ext):
                ma.build(); mb.build()
                x = sb.Isect.front_end(ma, mb, fa.data_ptr(), fb.data_ptr())
                res[0], res[1] = x.num_candidates, x.num_hits
                x.close()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(steps):
            with torch.cuda.stream(ext):
                l2_flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext); step(); e1.record(ext); e1.synchronize()
            tot += e0.elapsed_time(e1)
        ia, ib = fa.cpu().numpy(), fb.cpu().numpy()
        d = {"workload": desc, "tris_a": len(a[1]), "tris_b": len(b[1]), "ms_per_step": round(tot / steps, 4), "steps": steps,
             "candidate_pairs": int(res[0]), "intersecting_pairs": int(res[1]),
             "candidate_pairs_per_s": res[0] / (tot / steps * 1e-3), "inside_a": int(ia.sum()), "inside_b": int(ib.sum())}
        pin = pins.get(name)
        if pin:
            d["vs_reference_pins"] = {"counts": bool((res[0], res[1], int(ia.sum()), int(ib.sum())) ==
                                                     (pin["n_pairs"], pin["n_hits"], pin["inside_a_count"], pin["inside_b_count"])),
                                      "source": "tests/golden/fullsize.json (unmodified reference over all faces)"}
        out[name] = d
        ma.close(); mb.close()
    return out


def host_threads():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 32))
    except AttributeError:
        return max(1, min(os.cpu_count() or 1, 32))


def sample_description(detail, threads):
    if detail["classify_points"] >= detail["classify_points_total"]:
        return ("the whole workload, nothing extrapolated: prepare() x2 (2 threads) + broad phase + predicate loop + "
                "isPointInMesh on all %d face centroids x 3 axes over %d threads" % (detail["classify_points_total"], threads))
    return ("full prepare() x2 (2 threads) + full broad phase + full predicate loop; isPointInMesh on %d of the %d "
            "face centroids x 3 axes over %d threads, scaled to all faces" % (detail["classify_points"],
                                                                            detail["classify_points_total"], threads))


def config_dict(desc, name, nA, nB):
    """The same `config` object in both arms (the driver compares them)."""
    return {"workload": desc, "name": name, "tris_a": int(nA), "tris_b": int(nB),
            "l2": "512 MiB buffer written between timed steps (L2 flush)",
            "parallelism": "N > 1: the faces of both meshes dealt to the ranks by centroid z, every rank builds only what its slab can meet"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    a, b, desc = workload(args.config)
    threads = host_threads()
    times, H, detail = [], 0, None
    for i in range(args.warmup + args.steps):
        t, H, detail = cpu_reference_step(a, b, None, threads)   # every step is the whole workload, measured
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = H / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(desc, args.config, len(a[1]), len(b[1])),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": detail["kind"],
                         "sample": sample_description(detail, threads)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "front_end_ms": sec * 1e3, "detail": {k: (round(v, 6) if isinstance(v, float) else v) for k, v in detail.items() if not k.startswith("_")},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------

def shard_bounds(n, world):
    cuts = [((n * r // world) // 32) * 32 for r in range(world)] + [n]
    return cuts


def run_ours(args):
    import torch
    import torch.distributed as dist
    import solidboolean_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # Native libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner
    # there): stdout is kept for the ONE JSON line -- everything else of this process goes to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(sb.LIB_PATH):
        if local == 0:
            from solidboolean_b200.build import build
            build()
        if world > 1:
            dist.barrier()

    # clocks are sampled for the whole run (nvidia-smi needs a moment to start); the
    # samples falling into each timed region are reported
    sampler = ClockSampler(local)
    sampler.start()
    a, b, desc = workload(args.config)
    nA, nB = len(a[1]), len(b[1])
    nVA, nVB = len(a[0]), len(b[0])
    ctx = sb.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # pinned host copies of the inputs (what a caller of the C ABI hands over)
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (a[0], a[1].view(np.int32), b[0], b[1].view(np.int32))]

    cutsA, cutsB = shard_bounds(nA, world), shard_bounds(nB, world)
    a0, a1 = cutsA[rank], cutsA[rank + 1]
    b0, b1 = cutsB[rank], cutsB[rank + 1]

    dev = torch.device("cuda", local)
    # ONE buffer and ONE collective per step: [flags of A | flags of B | world x packed hit record].  Every byte
    # is written by exactly one rank (a face belongs to one rank, a record slot to its rank) and starts
    # out zero, so an all_reduce(sum) over int32 lanes IS the gather -- no carry ever crosses a byte.
    flag_bytes = (nA + nB + 15) // 16 * 16
    l2_flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    hit_cap = [4096]  # padded hit capacity per rank; grows to fit (a retry costs one extra exchange)
    xbuf = {}         # exchange buffer, re-made when the padding grows
    gathered = {}

    def make_xbuf():
        cap = hit_cap[0]
        rec = (16 + 8 * cap + 48 * cap + 15) // 16 * 16       # header + int32[cap,2] + float64[cap,6]
        buf = torch.zeros(flag_bytes + (world * rec if world > 1 else 0), dtype=torch.uint8, device=dev)
        xbuf.update(cap=cap, rec=rec, buf=buf, buf32=buf.view(torch.int32), checked=None)
        gathered["hits"] = buf[flag_bytes:]
    make_xbuf()

    def flags_views():
        buf = xbuf["buf"]
        return buf[:nA], buf[nA:nA + nB]
    flagsA, flagsB = flags_views()

    def gather_results(x):
        """The exchange of the shard results: sb_isect_pack_device writes this rank's record
        {nCand, nHit, hit pairs, segments} straight from the library's device buffers into its slot, and one
        NCCL all_reduce over the whole buffer hands every rank all flags and all records.  No host round
        trip in the steady state: the counts are only read back while the padding is still being sized."""
        while True:
            cap, rec = xbuf["cap"], xbuf["rec"]
            x.pack_device(xbuf["buf"].data_ptr() + flag_bytes + rank * rec, cap)
            dist.all_reduce(xbuf["buf32"])
            if xbuf["checked"] == cap:
                break
            cnt = xbuf["buf"][flag_bytes:].view(world, rec)[:, :16].contiguous().view(torch.int64).view(world, 2).cpu()
            if int(cnt[:, 1].max()) <= cap // 2:
                xbuf["checked"] = cap
                break
            # some rank is close to the padding: once more, with room (every rank takes the same decision)
            hit_cap[0] = int(cnt[:, 1].max()) * 3 + 64
            return "again"
        return None

    def gathered_counts():
        """Global (P, H, largest per-rank H) from the records of the last exchange (outside the timed region)."""
        rec = xbuf["rec"]
        cnt = xbuf["buf"][flag_bytes:].view(world, rec)[:, :16].contiguous().view(torch.int64).view(world, 2).cpu()
        return int(cnt[:, 0].sum()), int(cnt[:, 1].sum()), int(cnt[:, 1].max())

    # ---------------- resident loop: `value` ----------------
    ma = ctx.mesh(a[0], a[1], build=False)
    mb = ctx.mesh(b[0], b[1], build=False)
    ctx.synchronize()
    ctx.enable_timing(True)

    rays_cands = [0, 0]
    paths = [0, 0, 0, 0, 0]

    shard = sb.Shard(ma, mb, rank, world) if world > 1 else None

    def resident_step():
        # (one GPU: no torch call inside the step, so torch's current stream is left alone -- see e2e_step)
        with (torch.cuda.stream(ext) if world > 1 else contextlib.nullcontext()):
            while True:
                if world > 1:
                    # sb_shard_front_end: this rank's slab of faces -- selection, build of the selected triangles
                    # only, broad phase, predicate, both classifications; flags of its own faces into the zeroed buffer
                    xbuf["buf"].zero_()
                    fa, fb = flags_views()
                    x = shard.front_end(fa.data_ptr(), fb.data_ptr())
                else:
                    ma.build(); mb.build()          # concurrent: each mesh builds on its own stream
                    # sb_front_end: intersection on the context stream, the two
                    # classification directions overlapped on internal streams
                    x = sb.Isect.front_end(ma, mb, flagsA.data_ptr(), flagsB.data_ptr())
                rays_cands[0], rays_cands[1] = ctx.classify_stats()
                again = gather_results(x) if world > 1 else None
                P, H = x.num_candidates, x.num_hits
                paths[:] = x.path_counts()
                x.close()
                if again is None:
                    break
                make_xbuf()                          # the padding grew: same step once more
        return P, H

    def timed_loop(step_fn, steps, warmup, device_timed):
        for _ in range(warmup):
            res = step_fn()
        torch.cuda.synchronize()
        ctx.reset_timing()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = 0.0
        t_begin = time.monotonic()
        for _ in range(steps):
            with torch.cuda.stream(ext):
                l2_flush.zero_()          # evict L2 between timed iterations (untimed)
            torch.cuda.synchronize()
            if device_timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                res = step_fn()
                e1.record(ext)
                e1.synchronize()
                total_ms += e0.elapsed_time(e1)
            else:
                t0 = time.perf_counter()
                res = step_fn()
                torch.cuda.synchronize()
                total_ms += (time.perf_counter() - t0) * 1e3
        clocks = sampler.summary(t_begin, time.monotonic())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, res, clocks

    # `value`: the step as a caller runs it -- no stage events on the library's streams (sb_context_enable_timing is a
    # diagnostic: its two event records around every stage cost ~4 % of a step).  The stage breakdown and the
    # roofline's kernel time come from a second loop of the same K steps WITH the stage events (`ms_per_step_instrumented`).
    ctx.enable_timing(False)
    ms_step, (P, H), clocks = timed_loop(resident_step, args.steps, args.warmup, True)
    ctx.enable_timing(True)
    ms_step_instr, (P, H), _ = timed_loop(resident_step, args.steps, 1, True)
    if world > 1:
        P, H, hmax = gathered_counts()
        assert hmax <= hit_cap[0], "hit padding overflow in the exchange"
    P, H = int(P), int(H)
    stage_ms, launches = ctx.timing()
    stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
    stage_by_rank = None
    if world > 1:  # every rank's stage times (the step is as long as the slowest rank + the exchange)
        stage_by_rank = [None] * world
        dist.all_gather_object(stage_by_rank, {k: round(v, 4) for k, v in stage_ms.items()})
    launches_per_step = launches / args.steps
    flagsA, flagsB = flags_views()
    insideA, insideB = int(flagsA.sum().item()), int(flagsB.sum().item())
    mg_parity, shard_info = None, None
    if world > 1:
        shard_info = [None] * world
        dist.all_gather_object(shard_info, shard.info())
        if rank == 0:
            # the gathered bytes against the single-GPU output of the same run (outside the timed region)
            rec = xbuf["rec"]
            ev = gathered["hits"].view(world, rec).cpu().numpy()
            parts_ab, parts_seg = [], []
            for r in range(world):
                nh = int(ev[r, 8:16].view(np.int64)[0])
                parts_ab.append(ev[r, 16:16 + 8 * nh].view(np.uint32).reshape(-1, 2))
                parts_seg.append(ev[r, 16 + 8 * hit_cap[0]:16 + 8 * hit_cap[0] + 48 * nh].view(np.float64).reshape(-1, 6))
            gab, gseg = np.concatenate(parts_ab), np.concatenate(parts_seg)
            order = np.lexsort((gab[:, 1], gab[:, 0]))
            got_flags = xbuf["buf"][:nA + nB].cpu().numpy()
            with torch.cuda.stream(ext):
                ma.build(); mb.build()
                one = torch.zeros(nA + nB, dtype=torch.uint8, device=dev)
                x1 = sb.Isect.front_end(ma, mb, one.data_ptr(), one.data_ptr() + nA)
                h1, s1 = x1.hits()
                p1 = x1.num_candidates
                x1.close()
            mg_parity = {"what": "hit pairs + segments gathered from all ranks (sorted) and the summed per-face flags against the "
                                 "single-GPU front end run by rank 0 in the same process",
                         "candidate_pairs_equal": bool(p1 == P), "hit_pairs_identical": bool(np.array_equal(gab[order], h1)),
                         "segments_bit_identical": bool(gseg[order].tobytes() == s1.tobytes()),
                         "flags_identical": bool(np.array_equal(got_flags, one.cpu().numpy()))}
            del one
    rays, cands = rays_cands
    if world > 1:
        rc = torch.tensor([rays, cands], dtype=torch.int64, device=dev)
        dist.all_reduce(rc)
        rays_total, cands_total = int(rc[0]), int(rc[1])
    else:
        rays_total, cands_total = rays, cands

    # ---------------- host-buffer loop: `e2e` ----------------
    out_in_a_t = torch.zeros(nA, dtype=torch.uint8).pin_memory()
    out_in_b_t = torch.zeros(nB, dtype=torch.uint8).pin_memory()

    host_hits = {}   # pinned landing buffers of the e2e loop, grown on demand

    # N > 1: every rank uploads 1/N of the geometry bytes; the rest arrives over NVLink (one all_gather)
    geo_sizes = [24 * nVA, 24 * nVB, 12 * nA, 12 * nB]
    geo_off = [0, geo_sizes[0], geo_sizes[0] + geo_sizes[1], geo_sizes[0] + geo_sizes[1] + geo_sizes[2]]
    geo_total = sum(geo_sizes)
    geo_slice = ((geo_total + world - 1) // world + 255) // 256 * 256
    geo_host = geo_dev = None
    if world > 1:
        geo_host = torch.zeros(geo_slice * world, dtype=torch.uint8).pin_memory()
        for off, t in zip(geo_off, (pin[0], pin[2], pin[1], pin[3])):
            flat = t.view(-1).view(torch.uint8)
            geo_host[off:off + flat.numel()].copy_(flat)
        geo_dev = torch.empty(geo_slice * world, dtype=torch.uint8, device=dev)

    def host_front_end(m1, m2):
        # the reference-facing call with HOST buffers: flags and hits land in pinned memory (sb_front_end_host)
        if "cap" not in host_hits:
            cap = max(1024, int(H) * 3 // 2 + 64)
            host_hits.update(cap=cap, ab=torch.zeros(2 * cap, dtype=torch.int32).pin_memory(),
                             seg=torch.zeros(6 * cap, dtype=torch.float64).pin_memory())
        return sb.Isect.front_end_host(m1, m2, out_in_a_t.data_ptr(), out_in_b_t.data_ptr(), host_hits["ab"].data_ptr(),
                                       host_hits["seg"].data_ptr(), host_hits["cap"])

    def e2e_step():
        # (one GPU: the step makes no torch call -- everything goes through the C ABI -- so torch's current stream is left
        #  alone: entering / leaving torch.cuda.stream() costs ~40 us of host time per step)
        with (torch.cuda.stream(ext) if world > 1 else contextlib.nullcontext()):
            if world > 1:
                mine = geo_dev[rank * geo_slice:(rank + 1) * geo_slice]
                mine.copy_(geo_host[rank * geo_slice:(rank + 1) * geo_slice], non_blocking=True)   # H2D of this rank's share
                dist.all_gather_into_tensor(geo_dev, mine)
                base = geo_dev.data_ptr()
                # the step's geometry into the rank's two meshes (sb_mesh_update, device to device); the shard bound
                # to them sizes its launches from the previous step and lets the device confirm the plan
                ma.update(base + geo_off[0], base + geo_off[2], on_device=True)
                mb.update(base + geo_off[1], base + geo_off[3], on_device=True)
                xa = xb = sh = None
                xbuf["buf"].zero_()
                fa, fb = flags_views()
                x = shard.front_end(fa.data_ptr(), fb.data_ptr())
                assert gather_results(x) is None, "the hit padding was sized by the resident loop"
                Pg, Hg = x.num_candidates, x.num_hits
            elif e2e_mode[0] == "update":
                # host buffers in through the C ABI: the step's geometry goes into the two meshes the caller keeps
                # (sb_mesh_update: H2D copies of all four arrays, B's overlapping A's build), then sb_mesh_build for
                # each -- a full rebuild from the new bytes, reference lists verified against the new counts
                xa = xb = None
                ma.update(pin[0].data_ptr(), pin[1].data_ptr())
                mb.update(pin[2].data_ptr(), pin[3].data_ptr())
                ma.build(); mb.build()
                x = host_front_end(ma, mb)
                Pg, Hg = x.num_candidates, x.num_hits
            else:
                # the same with two NEW meshes per step (sb_mesh_upload + sb_mesh_build + sb_mesh_destroy): allocation,
                # stream / event creation and the first-build round trip that sizes the reference lists are in the time
                xa = sb.Mesh.from_pointers(ctx, pin[0].data_ptr(), nVA, pin[1].data_ptr(), nA, build=False, keep=pin)
                xb = sb.Mesh.from_pointers(ctx, pin[2].data_ptr(), nVB, pin[3].data_ptr(), nB, build=False, keep=pin)
                xa.build(); xb.build()
                sh = None
                x = host_front_end(xa, xb)
                Pg, Hg = x.num_candidates, x.num_hits
            if rank == 0 and world == 1:
                # sb_front_end_host has already put everything into the pinned host buffers (flags of both meshes, hit
                # pairs + segments), each copy enqueued behind the stream that produced it; a hit list longer than
                # the landing buffer is fetched with sb_isect_hits into a larger one
                if host_hits["cap"] < x.num_hits:
                    cap = x.num_hits * 3 // 2 + 64
                    host_hits.update(cap=cap, ab=torch.zeros(2 * cap, dtype=torch.int32).pin_memory(),
                                     seg=torch.zeros(6 * cap, dtype=torch.float64).pin_memory())
                    sb._check(x.lib.sb_isect_hits(x.h, host_hits["ab"].data_ptr(), host_hits["seg"].data_ptr()))
            elif rank == 0:  # results back to the host: per-face flags, hit pairs + segments -- all into PINNED
                # buffers (the C ABI takes any host pointer; pageable ones make the copies staged and slow),
                # flag copies enqueued first so that the one synchronisation inside sb_isect_hits covers all
                out_in_a_t.copy_(fa, non_blocking=True)
                out_in_b_t.copy_(fb, non_blocking=True)
                if host_hits.get("cap", -1) < x.num_hits:
                    cap = x.num_hits * 3 // 2 + 64
                    host_hits.update(cap=cap, ab=torch.zeros(2 * cap, dtype=torch.int32).pin_memory(),
                                     seg=torch.zeros(6 * cap, dtype=torch.float64).pin_memory())
                if world > 1:
                    # rank 0 reads every rank's record (hit pairs + segments) out of the gathered buffer
                    if host_hits.get("rec") != xbuf["rec"] * world:
                        host_hits.update(rec=xbuf["rec"] * world, all=torch.zeros(xbuf["rec"] * world, dtype=torch.uint8).pin_memory())
                    host_hits["all"].copy_(gathered["hits"], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                else:
                    sb._check(x.lib.sb_isect_hits(x.h, host_hits["ab"].data_ptr(), host_hits["seg"].data_ptr()))
            x.close()
            if xa is not None:
                xa.close(); xb.close()
        return Pg, Hg

    ctx.enable_timing(False)
    e2e_mode = ["fresh"]
    e2e_fresh_ms = None
    if world == 1:
        e2e_fresh_ms, _, _ = timed_loop(e2e_step, max(3, args.steps // 4), 2, False)
        e2e_mode[0] = "update"
    e2e_ms, (P2, H2), _ = timed_loop(e2e_step, max(3, args.steps // 2), 2, False)
    if world > 1:
        P2, H2, _ = gathered_counts()
    P2, H2 = int(P2), int(H2)
    sampler.stop()
    h2d = 24 * (nVA + nVB) + 12 * (nA + nB) if world == 1 else geo_slice   # per rank
    d2h = (nA + nB) + 56 * H + 64 if world == 1 else (nA + nB) + xbuf["rec"] * world

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # classification bytes = 25 Q + 104 nTarget + 108 C, both launches of a step
        # (N > 1: per GPU -- rank 0's own query shard and candidates against its own kernel time and ONE GPU's peak)
        q_local = (a1 - a0) + (b1 - b0) if world == 1 else (nA + nB) // world
        cls_bytes = 25 * q_local + 104 * (nA + nB) + 108 * (cands if world > 1 else cands_total)
        cls_ms = stage_ms["classify"]
        traffic = None
        try:  # DRAM bytes of the classification kernels of one step, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("config") == args.config and world == 1:
                traffic = tj["classify_dram_bytes_per_step"]
        except Exception:
            pass
        achieved = cls_bytes / (cls_ms * 1e-3) / 1e9 if cls_ms > 0 else 0.0
        build_bytes = 24 * (nVA + nVB) + 296 * (nA + nB)
        broad_bytes = 48 * nA + 104 * nB + 8 * P
        narrow_bytes = 8 * P + 24 * (nVA + nVB) + 12 * (nA + nB) + 56 * H

        def gbs(bytes_, ms):
            return round(bytes_ / (ms * 1e-3) / 1e9, 1) if ms > 0 else None
        line = {
            "metric": METRIC, "value": H / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(desc, args.config, nA, nB),
            **({"multi_gpu": {"scheme": "faces dealt by centroid z (sb_shard_*): every rank selects and builds only the triangles its "
                                        "slab can meet; one all_gather of the packed hit records + one all_reduce of the flag bytes",
                              "shards": shard_info, "parity_vs_single_gpu": mg_parity,
                              "e2e_upload": "each rank copies 1/N of the geometry bytes from pinned host memory, one all_gather over "
                                            "NVLink hands everybody the rest"}} if world > 1 else {}),
            "front_end_ms": ms_step,
            "ms_per_step_instrumented": ms_step_instr,
            "instrumentation": "value / ms_per_step: K steps without stage events on the library's streams; stage_ms, roofline and "
                               "gpu_launches: a second loop of K steps with them (ms_per_step_instrumented)",
            "candidate_pairs": P, "intersecting_pairs": H, "inside_a": insideA, "inside_b": insideB,
            "candidate_pairs_per_s": P / (ms_step * 1e-3),
            "triangles_per_s": (nA + nB) / (ms_step * 1e-3),
            "rays_per_s": rays_total / (stage_ms["classify"] * 1e-3) if stage_ms["classify"] > 0 else None,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            **({"stage_ms_by_rank": stage_by_rank} if stage_by_rank else {}),
            "stage_gbs_algorithmic": {"build": gbs(build_bytes, stage_ms["build"]), "broad": gbs(broad_bytes, stage_ms["broad"]),
                                      "narrow": gbs(narrow_bytes, stage_ms["narrow"]), "classify": gbs(cls_bytes, cls_ms)},
            "roofline": {"bound": "hbm", "kernel": "classification: classify2_kernel, both directions (2 launches/step)" + (", rank 0's shard" if world > 1 else ""), "achieved": round(achieved, 1),
                         "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": cls_bytes,
                         "kernel_ms_per_step": round(cls_ms, 4)},
            "e2e": {"value": H2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    **({"how": "sb_mesh_update (H2D of all four arrays from pinned host memory) + sb_mesh_build x2 + sb_front_end_host + "
                               "D2H of hit pairs, segments and per-face flags, every step; the two meshes are kept between steps",
                        "ms_per_step_new_meshes_every_step": e2e_fresh_ms} if e2e_fresh_ms is not None else {})},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        # FP64 view of the predicate stage: flops = 41 R1 + 82 R2 + 139 R3 + 185 H
        # (coplanar pairs counted like R3), against the measured non-FMA DMUL+DADD rate
        try:
            nofma, fma = ctx.fp64_peak()
            r1, r2, cop, r3, hh = paths
            flops = 41 * r1 + 82 * r2 + 139 * (r3 + cop) + 185 * hh
            pms = stage_ms.get("predicate", 0.0)
            line["fp64"] = {"predicate_ms_per_step": round(pms, 5), "pairs": int(sum(paths)),
                            "exit_histogram": {"plane2_reject": r1, "plane1_reject": r2, "coplanar": cop,
                                               "interval_reject": r3, "segment": hh},
                            "flops_per_step": flops,
                            "achieved_gflops": round(flops / (pms * 1e-3) / 1e9, 2) if pms > 0 else None,
                            "peak_nofma_gflops": round(nofma, 1), "peak_fma_gflops": round(fma, 1),
                            "frac_of_nofma_peak": round(flops / (pms * 1e-3) / 1e9 / nofma, 5) if pms > 0 and nofma > 0 else None,
                            "note": "shard of rank 0" if world > 1 else "whole workload"}
        except Exception as e:  # the microbenchmark must never break the bench line
            line["fp64"] = {"error": str(e)}
        if world == 1 and args.config == "c3":
            # the other single-GPU configurations of BASELINE.json, measured in the same run with the same step
            # (parity-test cases, not the bench line): C2 and C4 at its stated size, results checked against the
            # hashes the unmodified reference left in tests/golden/fullsize.json
            try:
                line["other_configs"] = other_configs(sb, ctx, ext, l2_flush, torch)
            except Exception as e:
                line["other_configs"] = {"error": str(e)}
        if world == 1:
            try:
                line["next_rows"] = next_rows(sb, ctx, ma, mb, a, b, not args.no_cpu_baseline)
            except Exception as e:  # evidence for the widened rows, never allowed to break the bench line
                line["next_rows"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sec, Href, detail = cpu_reference_step(a, b, None, threads)
            line["cpu_baseline"] = {"value": Href / sec, "unit": UNIT, "cores": threads, "kind": detail["kind"],
                                    "sample": sample_description(detail, threads), "front_end_ms": sec * 1e3,
                                    "measured_ms": detail["measured_s"] * 1e3, "extrapolated_ms": detail["extrapolated_s"] * 1e3,
                                    "H": Href, "P": detail["P"], "inside_a": detail["inside_a"], "inside_b": detail["inside_b"]}
            # the whole workload was run on the CPU: compare everything the e2e step brought back with it
            try:
                fa, fb = detail["_flags"]
                rab, rseg = detail["_hits"]
                order = np.lexsort((rab[:, 1], rab[:, 0]))
                gab = host_hits["ab"][:2 * H2].numpy().reshape(-1, 2).astype(np.uint32)
                gseg = host_hits["seg"][:6 * H2].numpy().reshape(-1, 6)
                line["parity_vs_cpu"] = {
                    "what": "results of the last host-buffer step against the CPU baseline run above (all faces, all pairs)",
                    "candidate_pairs_equal": bool(P2 == detail["P"]),
                    "hit_pairs_identical": bool(np.array_equal(gab, rab[order].astype(np.uint32))),
                    "segments_bit_identical": bool(gseg.tobytes() == np.ascontiguousarray(rseg[order]).tobytes()),
                    "flags_a_identical": bool(np.array_equal(out_in_a_t.numpy(), fa)),
                    "flags_b_identical": bool(np.array_equal(out_in_b_t.numpy(), fb))}
            except Exception as e:
                line["parity_vs_cpu"] = {"error": str(e)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    # Orderly teardown, then a NORMAL interpreter exit (exit hooks must run: the driver records the
    # shared objects this process mapped from one).  Everything that was used on the library's stream
    # is released while that stream is alive -- torch's caching allocators record events on the stream
    # a block was used on -- and only then is the context destroyed.
    torch.cuda.synchronize()
    gathered.clear()
    host_hits.clear()
    xbuf.clear()
    if shard is not None:
        shard.close()
    del geo_host, geo_dev
    del out_in_a_t, out_in_b_t, flagsA, flagsB, l2_flush, pin
    ma.close(); mb.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    del ext
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()       # pinned-host cache: its blocks remember the library's stream
    except Exception:
        pass
    torch.cuda.synchronize()
    ctx.close()
    sys.stdout.flush()
    sys.stderr.flush()


# ----------------------------------------------------------------------------
# BASELINE configs[4] ("C5"): a batch of 1,000 small booleans (icosphere k=4 pairs, 5,120 + 5,120
# triangles, seeded offsets), jobs dealt job_id mod N over the GPUs, no collective.
# On every rank the jobs form ONE batch mesh per side (sb_batch_upload): one sort, one leaf / grid
# build and one front-end launch sequence for all of them.

C5_DESC = "batch of %d boolean front ends, icosphere k=4 pairs (5,120 + 5,120 triangles each, seeded offsets), BASELINE configs[4]"


def c5_jobs(n_jobs, rank=0, world=1):
    from solidboolean_b200 import meshgen
    ids = list(range(rank, n_jobs, world))
    return ids, [meshgen.config_c5_job(j) for j in ids]


def c5_reference_jobs(jobs, threads):
    """The reference (one SolidBoolean front end per job: prepare x2, search, predicate loop,
    isPointInMesh on every face centroid) over all host cores, one job per thread at a time.
    -> (seconds, total hits, per-job results)."""
    from oracle import Oracle, Ref
    use_ref = Ref.available()
    out = [None] * len(jobs)

    def one(k):
        a, b = jobs[k]
        if use_ref:
            R = Ref.get()
            ma, mb = R.mesh(*a), R.mesh(*b)
            op = R.op(ma, mb)
            pairs = op.search()
            ret, cop, hit, seg, _ = op.predicate(pairs)
            fa = op.classify(1, ma.centroids())[0]
            fb = op.classify(0, mb.centroids())[0]
            op.close(); ma.close(); mb.close()
        else:
            O = Oracle.get()
            pairs = O.candidate_pairs(a, b)
            ret, cop, hit, seg = O.predicate_pairs(a, b, pairs)
            fa = O.classify(b, O.centroids(*a))[0]
            fb = O.classify(a, O.centroids(*b))[0]
        h = hit.astype(bool)
        order = np.lexsort((pairs[h][:, 1], pairs[h][:, 0]))
        out[k] = (len(pairs), pairs[h][order], seg[h][order], fa, fb)

    def worker(t):
        for k in range(t, len(jobs), threads):
            one(k)
    t0 = time.perf_counter()
    ws = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    [w.start() for w in ws]
    [w.join() for w in ws]
    sec = time.perf_counter() - t0
    return sec, sum(len(r[1]) for r in out), out, ("reference" if use_ref else "port")


def c5_config(n_jobs):
    return {"workload": C5_DESC % n_jobs, "name": "c5", "jobs": n_jobs, "tris_a": 5120 * n_jobs, "tris_b": 5120 * n_jobs,
            "l2": "512 MiB buffer written between timed steps (L2 flush)",
            "parallelism": "jobs dealt job_id mod N over the ranks, one batch mesh per side and rank, no collective"}


def run_reference_c5(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = host_threads()
    # a bounded sample per step: 4 jobs per host thread, scaled to the whole batch by the job count
    sample = min(args.jobs, 4 * threads)
    _, jobs = c5_jobs(sample)
    times, H = [], 0
    for i in range(args.warmup + args.steps):
        sec, H, _, kind = c5_reference_jobs(jobs, threads)
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times)) * args.jobs / sample
    Hs = H * args.jobs / sample
    val = Hs / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": c5_config(args.jobs),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d of the %d jobs per step (4 per host thread), each job the reference's whole front end; "
                                       "time and hits scaled by the job count" % (sample, args.jobs)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_job": sec * 1e3 / args.jobs, "jobs_per_s": args.jobs / sec}
    print(json.dumps(line), flush=True)


def run_c5(args):
    import torch
    import torch.distributed as dist
    import solidboolean_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    sampler.start()
    ids, jobs = c5_jobs(args.jobs, rank, world)
    J = len(jobs)
    ctx = sb.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    def work():
        # the batch as a caller hands it over: concatenated arrays + start offsets, in pinned host memory
        def cat(side):
            xs = [j[side][0] for j in jobs]
            ts = [j[side][1] for j in jobs]
            return (np.concatenate(xs), np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.uint64),
                    np.concatenate(ts), np.concatenate([[0], np.cumsum([len(t) for t in ts])]).astype(np.uint64))
        hostA, hostB = cat(0), cat(1)
        pitch = sb.batch_pitch(hostA[0], hostB[0])
        if world > 1:   # one pitch for the whole run (any rank's value would do for its own batches; kept equal for the record)
            t = torch.tensor([pitch], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pitch = float(t.item())
        pinned = [[torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (h[0], h[2].view(np.int32))] for h in (hostA, hostB)]
        nA, nB = len(hostA[2]), len(hostB[2])
        nVA, nVB = len(hostA[0]), len(hostB[0])
        flagsA = torch.zeros(nA, dtype=torch.uint8, device=dev)
        flagsB = torch.zeros(nB, dtype=torch.uint8, device=dev)
        l2_flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

        def make(side, build):
            h, pin = (hostA, hostB)[side], pinned[side]
            return sb.Mesh.batch_arrays(ctx, pin[0].numpy(), h[1], pin[1].numpy().view(np.uint32), h[3], pitch, build=build)
        A, B = make(0, False), make(1, False)
        ctx.synchronize()
        ctx.enable_timing(True)
        stats = [0, 0]

        def resident_step():
            with torch.cuda.stream(ext):
                A.build(); B.build()
                x = sb.Isect.front_end(A, B, flagsA.data_ptr(), flagsB.data_ptr())
                stats[0], stats[1] = ctx.classify_stats()
                P, H = x.num_candidates, x.num_hits
                x.close()
            return P, H

        def timed_loop(step_fn, steps, warmup, device_timed):
            for _ in range(warmup):
                res = step_fn()
            torch.cuda.synchronize()
            ctx.reset_timing()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            total_ms = 0.0
            t_begin = time.monotonic()
            for _ in range(steps):
                with torch.cuda.stream(ext):
                    l2_flush.zero_()
                torch.cuda.synchronize()
                if device_timed:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ext)
                    res = step_fn()
                    e1.record(ext)
                    e1.synchronize()
                    total_ms += e0.elapsed_time(e1)
                else:
                    t0 = time.perf_counter()
                    res = step_fn()
                    torch.cuda.synchronize()
                    total_ms += (time.perf_counter() - t0) * 1e3
            clocks = sampler.summary(t_begin, time.monotonic())
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) / steps, res, clocks

        ctx.enable_timing(False)      # (see the C3 leg: `value` without stage events, the breakdown from a second loop with them)
        ms_step, (P, H), clocks = timed_loop(resident_step, args.steps, args.warmup, True)
        ctx.enable_timing(True)
        ms_step_instr, (P, H), _ = timed_loop(resident_step, args.steps, 1, True)
        stage_ms, launches = ctx.timing()
        stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
        rays, cands = stats

        # host-buffer loop: upload of the rank's whole batch, builds, front end, everything back to the host
        out_a = torch.zeros(nA, dtype=torch.uint8).pin_memory()
        out_b = torch.zeros(nB, dtype=torch.uint8).pin_memory()
        host_hits = {}
        last = {}

        def e2e_step():
            with torch.cuda.stream(ext):
                xa, xb = make(0, False), make(1, False)
                xa.build(); xb.build()
                if "cap" not in host_hits:
                    cap = max(1024, int(H) * 3 // 2 + 64)
                    host_hits.update(cap=cap, ab=torch.zeros(2 * cap, dtype=torch.int32).pin_memory(),
                                     seg=torch.zeros(6 * cap, dtype=torch.float64).pin_memory())
                # flags of both batches and the hit list land in pinned host memory (sb_front_end_host)
                x = sb.Isect.front_end_host(xa, xb, out_a.data_ptr(), out_b.data_ptr(), host_hits["ab"].data_ptr(),
                                            host_hits["seg"].data_ptr(), host_hits["cap"])
                if host_hits["cap"] < x.num_hits:
                    cap = x.num_hits * 3 // 2 + 64
                    host_hits.update(cap=cap, ab=torch.zeros(2 * cap, dtype=torch.int32).pin_memory(),
                                     seg=torch.zeros(6 * cap, dtype=torch.float64).pin_memory())
                    sb._check(x.lib.sb_isect_hits(x.h, host_hits["ab"].data_ptr(), host_hits["seg"].data_ptr()))
                last["ranges"] = x.job_ranges()
                res = (x.num_candidates, x.num_hits)
                x.close(); xa.close(); xb.close()
            return res

        ctx.enable_timing(False)
        e2e_ms, (P2, H2), _ = timed_loop(e2e_step, max(3, args.steps // 2), 2, False)
        sampler.stop()
        tot = torch.tensor([P, H, P2, H2, rays, cands, J], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        Pg, Hg, P2g, H2g, rays_g, cands_g, Jg = [int(v) for v in tot.tolist()]

        # parity: jobs of this rank against the CPU (the reference where it was built, else the port), all fields
        want = max(1, (64 + world - 1) // world)
        sample = list(range(0, J, max(1, J // want)))[:want] if not args.no_cpu_baseline else []
        ok, checked = True, 0
        if sample:
            _, _, refs, kind = c5_reference_jobs([jobs[k] for k in sample], host_threads())
            r = last["ranges"]
            gab = host_hits["ab"][:2 * H2].numpy().reshape(-1, 2).astype(np.int64)
            gseg = host_hits["seg"][:6 * H2].numpy().reshape(-1, 6)
            ta, tb = hostA[3].astype(np.int64), hostB[3].astype(np.int64)
            fa, fb = out_a.numpy(), out_b.numpy()
            for k, ref in zip(sample, refs):
                mine = gab[r[k]:r[k + 1]] - [ta[k], tb[k]]
                ok = ok and np.array_equal(mine, ref[1].astype(np.int64)) and gseg[r[k]:r[k + 1]].tobytes() == ref[2].tobytes() \
                    and np.array_equal(fa[ta[k]:ta[k + 1]], ref[3]) and np.array_equal(fb[tb[k]:tb[k + 1]], ref[4])
                checked += 1
        par = torch.tensor([1 if ok else 0, checked], dtype=torch.int64, device=dev)
        if world > 1:
            okt = par[:1].clone()
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            dist.all_reduce(par[1:])
            par[0] = okt[0]
        if rank == 0:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
            cls_bytes = 25 * (nA + nB) + 104 * (nA + nB) + 108 * cands    # SURVEY 8d, rank 0's batch
            cls_ms = stage_ms["classify"]
            achieved = cls_bytes / (cls_ms * 1e-3) / 1e9 if cls_ms > 0 else 0.0
            n_jobs = Jg
            line = {
                "metric": METRIC, "value": Hg / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": c5_config(n_jobs),
                "jobs": n_jobs, "ms_per_job": ms_step / n_jobs, "jobs_per_s": n_jobs / (ms_step * 1e-3),
                "ms_per_step_instrumented": ms_step_instr,
                "instrumentation": "value / ms_per_step: K steps without stage events on the library's streams; stage_ms, roofline and "
                                   "gpu_launches: a second loop of K steps with them (ms_per_step_instrumented)",
                "candidate_pairs": Pg, "intersecting_pairs": Hg,
                "triangles_per_s": 10240 * n_jobs / (ms_step * 1e-3),
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "roofline": {"bound": "hbm", "kernel": "classification of rank 0's batch: classify2_kernel, both directions",
                             "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                             "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                             "algorithmic_bytes_per_step": cls_bytes, "kernel_ms_per_step": round(cls_ms, 4)},
                "e2e": {"value": H2g / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "ms_per_job": e2e_ms / n_jobs,
                        "h2d_bytes_per_step": 24 * (nVA + nVB) + 12 * (nA + nB) + 16 * (J + 1) * 2,
                        "d2h_bytes_per_step": (nA + nB) + 56 * int(H2) + 64, "bytes_are": "rank 0's share"},
                "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
                "parity": {"jobs_checked": int(par[1]), "identical": bool(int(par[0])),
                           "what": "hit pairs, segments (bit for bit) and per-face flags of the sampled jobs against the CPU, from the last host-buffer step"},
                "clocks": clocks,
            }
            if world == 1 and not args.no_cpu_baseline:
                threads = host_threads()
                ns = min(J, 4 * threads)
                sec, Hs, _, kind = c5_reference_jobs(jobs[:ns], threads)
                line["cpu_baseline"] = {"value": Hs / sec, "unit": UNIT, "cores": threads, "kind": kind,
                                        "sample": "%d of the %d jobs (4 per host thread), each the reference's whole front end" % (ns, n_jobs),
                                        "ms_per_job": sec * 1e3 / ns, "jobs_per_s": ns / sec}
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        torch.cuda.synchronize()
        A.close(); B.close()

    # everything that lives in pinned or device memory is a local of work(): it is gone (and, after the
    # collection below, really freed) before the caches are emptied and the library's streams destroyed
    work()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    del ext
    gc.collect()
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()
    except Exception:
        pass
    torch.cuda.synchronize()
    ctx.close()
    sys.stdout.flush()
    sys.stderr.flush()


def next_rows(sb, ctx, ma, mb, a, b, with_cpu):
    """SURVEY 8f rows 1-3, outside the timed step: the per-triangle intersection contexts (the
    pair-loop body of combine(), :296-339), uncut triangles + half-edge map
    (addUnintersectedTriangles, reference src/solidboolean.cpp:250-286) and the face groups of the
    uncut triangles (the flood of buildFaceGroups, :229-238) on the device, device time from the
    library's stage events; beside them the reference's own two functions on one host core."""
    ctx.enable_timing(True)
    ma.build(); mb.build()
    x = ma.intersect(mb)
    best = None
    for _ in range(4):
        ctx.reset_timing()
        ua = x.uncut(0, 0, 0)
        ub = x.uncut(1, len(a[0]), ua.num_triangles)
        he_ms = ctx.timing()[0]["halfedge"]
        ctx.reset_timing()
        ga, gb = ua.components()[1], ub.components()[1]
        cc_ms = ctx.timing()[0]["halfedge"]
        rec = {"halfedge_ms": round(he_ms, 4), "components_ms": round(cc_ms, 4),
               "uncut_triangles": [ua.num_triangles, ub.num_triangles], "face_groups": [ga, gb], "ok": [ua.ok, ub.ok]}
        if best is None or he_ms + cc_ms < best["halfedge_ms"] + best["components_ms"]:
            best = rec
        keep = (ua.half_edges(), ub.half_edges(), ua.components()[0], ub.components()[0]) if with_cpu and _ == 3 else None
        ua.close(); ub.close()
        # row 1: the per-triangle intersection contexts of both meshes (the pair-loop body of combine())
        ctx.reset_timing()
        cuts = [x.contexts(w) for w in (0, 1)]
        rec["contexts_ms"] = round(ctx.timing()[0]["contexts"], 4)
        rec["contexts"] = [len(c["tri"]) for c in cuts]
        if best is not rec and rec["contexts_ms"] < best.get("contexts_ms", 1e9):
            best["contexts_ms"], best["contexts"] = rec["contexts_ms"], rec["contexts"]
    fa, fb = x.face_flags()
    x_hits = x.hits()
    x.close()
    ctx.enable_timing(False)

    def bits_for(n):
        k = 1
        while (1 << k) < n:
            k += 1
        return k
    total = 0
    for m, voff, n in ((a, 0, best["uncut_triangles"][0]), (b, len(a[0]), best["uncut_triangles"][1])):
        passes = (2 * bits_for(voff + len(m[0])) + 7) // 8
        total += 13 * len(m[1]) + 16 * n + 36 * n + 72 * n * passes + 84 * n
    best["halfedge_algorithmic_bytes"] = total
    best["halfedge_gbs_algorithmic"] = round(total / (best["halfedge_ms"] * 1e-3) / 1e9, 1) if best["halfedge_ms"] > 0 else None
    if with_cpu:
        from oracle import Ref
        if Ref.available():
            R = Ref.get()
            op = R.op(R.mesh(*a), R.mesh(*b))
            (ra, rb), _t = op.uncut(fa, fb)
            la, _ga = op.uncut_groups(0, best["uncut_triangles"][0])
            lb, _gb = op.uncut_groups(1, best["uncut_triangles"][1])
            (ka, oa), (kb, ob), ca, cb = keep
            from oracle import Oracle
            hab, hseg = x_hits
            t0 = time.perf_counter()
            oc = [Oracle.get().cut_contexts(hab, hseg, w) for w in (0, 1)]
            port_ms = (time.perf_counter() - t0) * 1e3
            same_ctx = all(np.array_equal(d[k], o[k]) for d, o in zip(cuts, oc) for k in ("tri", "point_start", "edge_start", "edges")) \
                and all(d["points"].tobytes() == o["points"].tobytes() for d, o in zip(cuts, oc))
            comb = op.combine()     # the reference's own time-points around its pair loop (predicate included; it may fail later on)
            best["reference_cpu"] = {
                "pair_loop_stage_ms": round(float(comb["stage_ms"][1]), 1), "contexts_oracle_port_ms": round(port_ms, 2),
                "contexts_identical_to_oracle": bool(same_ctx),
                "cores": 1, "addUnintersectedTriangles_ms": round(op.uncut_ms(0, 0) + op.uncut_ms(0, 1), 1),
                "buildFaceGroups_ms": round(op.uncut_ms(1, 0) + op.uncut_ms(1, 1), 1),
                "identical": bool(np.array_equal(ra["keys"], ka) and np.array_equal(ra["owner"], oa)
                                  and np.array_equal(rb["keys"], kb) and np.array_equal(rb["owner"], ob)
                                  and np.array_equal(la, ca) and np.array_equal(lb, cb))}
            op.close()
    return best


def _as_tensor(torch, ptr, shape, dtype, dev):
    """Zero-copy torch view of a device pointer owned by the library."""
    n = int(np.prod(shape))
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Holder:
        pass
    h = _Holder()
    typestr = {torch.int32: "<i4", torch.float64: "<f8", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                  "strides": None}
    assert n * itemsize >= 0
    return torch.as_tensor(h, device=dev)


if __name__ == "__main__":
    args = parse_args()
    if args.config == "c5":
        (run_reference_c5 if args.impl == "reference" else run_c5)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
