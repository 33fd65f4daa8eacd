"""Multi-GPU shards (sb_shard_*) on ONE GPU: the n ranks of a front end are run one after
the other into the same flag arrays; the union of their hit lists and the summed flags must be the
single-GPU result, which the other tests pin to the oracle.  (The NCCL exchange itself is covered by
bench.py --gpus N, which compares the gathered bytes with the 1-GPU output inside the run.)"""
import numpy as np
import pytest

import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def run_ranks(ctx, a, b, n):
    import torch
    A, B = ctx.mesh(*a, build=False), ctx.mesh(*b, build=False)
    fa = torch.zeros(max(len(a[1]), 1), dtype=torch.uint8, device="cuda")
    fb = torch.zeros(max(len(b[1]), 1), dtype=torch.uint8, device="cuda")
    hits, segs, P, infos = [], [], 0, []
    for r in range(n):
        s = sb.Shard(A, B, r, n)
        for rep in range(2):       # the second call reuses the selections (rebuild path)
            x = s.front_end(fa.data_ptr(), fb.data_ptr())
            h, sg = x.hits()
            cnt = x.num_candidates
            x.close()
        hits.append(h); segs.append(sg); P += cnt
        infos.append(s.info())
        s.close()
    ctx.synchronize()
    hab = np.concatenate(hits) if hits else np.zeros((0, 2), np.uint32)
    seg = np.concatenate(segs) if segs else np.zeros((0, 6))
    order = np.lexsort((hab[:, 1], hab[:, 0]))
    out = (P, hab[order], seg[order], fa.cpu().numpy()[:len(a[1])], fb.cpu().numpy()[:len(b[1])], infos)
    A.close(); B.close()
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 8])
def test_shards_union_equals_single_gpu(ctx, oracle, n):
    a = meshgen.icosphere(5)
    b = meshgen.torus(128, 64, center=(0.013, 0.007, 0.011))
    P, hab, seg, fa, fb, infos = run_ranks(ctx, a, b, n)
    ref = oracle.candidate_pairs(a, b)
    ret, cop, hit, rseg = oracle.predicate_pairs(a, b, ref)
    h = hit.astype(bool)
    assert P == len(ref), "the ranks' candidate counts do not add up to the reference's"
    assert np.array_equal(hab, ref[h]) and seg.tobytes() == rseg[h].tobytes()
    assert np.array_equal(fa, oracle.classify(b, oracle.centroids(*a))[0])
    assert np.array_equal(fb, oracle.classify(a, oracle.centroids(*b))[0])
    if n > 1:   # the selections really are parts of the meshes
        assert max(i["selected_a"] for i in infos) < len(a[1]) and max(i["selected_b"] for i in infos) < len(b[1])
    assert all(i["fallbacks"] == 0 for i in infos)


def test_shards_third_ray_fallback(ctx, oracle):
    """Query faces whose centroids lie exactly on faces / vertices of the other mesh make the first two
    votes disagree: the third ray (along z) needs the whole target -- the fallback path."""
    a = meshgen.slab(8, 1.0, 0.3)
    b = meshgen.slab(8, 1.0, 0.3, center=(0.25, 0.125, 0.3))   # b's bottom face = a's top plane
    P, hab, seg, fa, fb, infos = run_ranks(ctx, a, b, 3)
    oa, pa, _ = oracle.classify(b, oracle.centroids(*a))
    ob, pb, _ = oracle.classify(a, oracle.centroids(*b))
    assert np.array_equal(fa, oa) and np.array_equal(fb, ob)
    ref = oracle.candidate_pairs(a, b)
    ret, cop, hit, rseg = oracle.predicate_pairs(a, b, ref)
    assert P == len(ref) and np.array_equal(hab, ref[hit.astype(bool)])
    if ((pa[:, 0] != pa[:, 1]).any() or (pb[:, 0] != pb[:, 1]).any()):
        assert sum(i["fallbacks"] for i in infos) > 0


def test_shards_degenerate_inputs(ctx, oracle):
    """Flat meshes (every centroid in one bin), more ranks than faces, an empty mesh."""
    import torch
    a = meshgen.slab(2, 1.0, 0.0)     # zero thickness: all z equal
    b = meshgen.icosphere(1, radius=0.4)
    P, hab, seg, fa, fb, infos = run_ranks(ctx, a, b, 4)
    ref = oracle.candidate_pairs(a, b)
    ret, cop, hit, rseg = oracle.predicate_pairs(a, b, ref)
    assert P == len(ref) and np.array_equal(hab, ref[hit.astype(bool)])
    assert np.array_equal(fa, oracle.classify(b, oracle.centroids(*a))[0])
    assert np.array_equal(fb, oracle.classify(a, oracle.centroids(*b))[0])
    e = (np.zeros((3, 3)), np.zeros((0, 3), np.uint32))
    P, hab, seg, fa, fb, infos = run_ranks(ctx, b, e, 2)
    assert P == 0 and len(hab) == 0 and not fa.any()
