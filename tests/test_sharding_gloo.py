"""CPU, world_size 2, gloo: the host-side sharding / gather logic of the
multi-GPU path without a GPU.  Each rank computes its shard with the
ORACLE standing in for the device kernels (this is a test of the partitioning and
the collectives, not of the kernels), results are exchanged exactly as
bench.py::gather_results does, and every rank must end up with the unsharded answer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_bounds
    from oracle import Oracle
    from solidboolean_b200 import meshgen
    O = Oracle.get()
    a = meshgen.icosphere(3)
    b = meshgen.torus(24, 12, center=(0.013, 0.007, 0.011))
    nA, nB = len(a[1]), len(b[1])
    # a Morton-like order stand-in: any fixed permutation known to all ranks
    order_a = np.random.default_rng(1).permutation(nA)
    order_b = np.random.default_rng(2).permutation(nB)
    cuts_a, cuts_b = shard_bounds(nA, world), shard_bounds(nB, world)
    assert cuts_a[0] == 0 and cuts_a[-1] == nA and all(c % 32 == 0 for c in cuts_a[:-1])
    mine_a = np.sort(order_a[cuts_a[rank]:cuts_a[rank + 1]])
    mine_b = np.sort(order_b[cuts_b[rank]:cuts_b[rank + 1]])
    # shard of the intersection: A's triangles in my range against all of B
    sub = (a[0], a[1][mine_a])
    pairs = O.candidate_pairs(sub, b)
    pairs[:, 0] = mine_a[pairs[:, 0]]
    ret, cop, hit, seg = O.predicate_pairs(a, b, pairs)
    hab, hseg = pairs[hit.astype(bool)].astype(np.int32), seg[hit.astype(bool)]
    # shard of the classification
    flags_a = torch.zeros(nA, dtype=torch.uint8)
    flags_b = torch.zeros(nB, dtype=torch.uint8)
    ia, _, _ = O.classify(b, O.centroids(*a)[mine_a])
    ib, _, _ = O.classify(a, O.centroids(*b)[mine_b])
    flags_a[torch.from_numpy(mine_a)] = torch.from_numpy(ia)
    flags_b[torch.from_numpy(mine_b)] = torch.from_numpy(ib)
    # ---- the exchange step (same protocol as bench.py::gather_results): one all_gather of
    # a packed per-rank record {nCand, nHit, pairs, segments}, one all_reduce over both flag arrays ----
    # (record layout = what sb_isect_pack_device writes; the flag bytes are reduced on int32 lanes over a
    # buffer padded to whole words: every byte has one owner, so no carry crosses a byte)
    flags_ab = torch.zeros((nA + nB + 15) // 16 * 16, dtype=torch.uint8)
    flags_ab[:nA] = flags_a
    flags_ab[nA:nA + nB] = flags_b
    cap = 8                                   # deliberately too small: exercises the regrow-and-repeat path
    while True:
        rec = 16 + 8 * cap + 48 * cap
        mine = torch.zeros(rec, dtype=torch.uint8)
        mine[:16].view(torch.int64).copy_(torch.tensor([len(pairs), len(hab)], dtype=torch.int64))
        n = min(len(hab), cap)
        if n:
            mine[16:16 + 8 * n].view(torch.int32).copy_(torch.from_numpy(hab[:n].reshape(-1)))
            mine[16 + 8 * cap:16 + 8 * cap + 48 * n].view(torch.float64).copy_(torch.from_numpy(hseg[:n].reshape(-1)))
        everyone = torch.empty(world * rec, dtype=torch.uint8)
        dist.all_gather_into_tensor(everyone, mine)
        counts_all = everyone.view(world, rec)[:, :16].contiguous().view(torch.int64).view(world, 2)
        if int(counts_all[:, 1].max()) <= cap:
            break
        cap = int(counts_all[:, 1].max()) * 3 // 2 + 64
    dist.all_reduce(flags_ab.view(torch.int32))
    flags_a, flags_b = flags_ab[:nA], flags_ab[nA:nA + nB]
    ev = everyone.view(world, rec)
    gab = np.concatenate([ev[r, 16:16 + 8 * int(counts_all[r, 1])].contiguous().view(torch.int32).view(-1, 2).numpy()
                          for r in range(world)])
    gseg = np.concatenate([ev[r, 16 + 8 * cap:16 + 8 * cap + 48 * int(counts_all[r, 1])].contiguous().view(torch.float64).view(-1, 6).numpy()
                           for r in range(world)])
    o = np.lexsort((gab[:, 1], gab[:, 0]))
    gab, gseg = gab[o], gseg[o]
    # ---- every rank compares with the unsharded oracle ----
    full = O.candidate_pairs(a, b)
    fret, fcop, fhit, fseg = O.predicate_pairs(a, b, full)
    ok = int(counts_all[:, 0].sum()) == len(full)
    ok &= np.array_equal(gab, full[fhit.astype(bool)].astype(np.int32))
    ok &= gseg.tobytes() == fseg[fhit.astype(bool)].tobytes()
    fa, _, _ = O.classify(b, O.centroids(*a))
    fb, _, _ = O.classify(a, O.centroids(*b))
    ok &= np.array_equal(flags_a.numpy(), fa) and np.array_equal(flags_b.numpy(), fb)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_unsharded():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_bounds_properties():
    sys.path.insert(0, ROOT)
    from bench import shard_bounds
    for n in (0, 1, 31, 32, 33, 1000, 1310720):
        for w in (1, 2, 4, 8):
            c = shard_bounds(n, w)
            assert c[0] == 0 and c[-1] == n and len(c) == w + 1
            assert all(c[i] <= c[i + 1] for i in range(w))
            assert all(x % 32 == 0 for x in c[:-1])
