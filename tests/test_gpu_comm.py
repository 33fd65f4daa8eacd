"""sb_comm_*: several GPUs driven from ONE host process behind the C ABI.
On a one-GPU box the ranks share device 0 (the same code path: one context, one shard, one host thread per
rank, peer copies of the flag bytes, OR on rank 0, merged hit lists); with more GPUs they spread out.
The gathered result must equal the single-GPU front end / the oracle byte for byte."""
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


def _devices(n):
    import torch
    g = torch.cuda.device_count()
    return [r % g for r in range(n)]


@pytest.mark.parametrize("n", [1, 2, 3])
def test_comm_equals_oracle(oracle, n):
    a = meshgen.icosphere(5)
    b = meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    c = sb.Comm(_devices(n))
    assert c.size == n
    c.set_meshes(a, b)
    ia, ib = c.front_end()
    ref = oracle.candidate_pairs(a, b)
    ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
    assert c.num_candidates == len(ref) and c.num_hits == int(hit.sum())
    hab, hseg = c.hits()
    assert np.array_equal(hab, ref[hit.astype(bool)]) and hseg.tobytes() == seg[hit.astype(bool)].tobytes()
    oa, _, _ = oracle.classify(b, oracle.centroids(*a))
    ob, _, _ = oracle.classify(a, oracle.centroids(*b))
    assert np.array_equal(ia, oa) and np.array_equal(ib, ob)
    infos = [c.rank_info(r) for r in range(n)]
    assert sum(i["hits"] for i in infos) == c.num_hits and sum(i["candidates"] for i in infos) == c.num_candidates
    # the next frame of the same sizes: moved geometry through the same meshes and shards
    b2 = (b[0] + np.array([0.05, -0.02, 0.03]), b[1])
    c.set_meshes(a, b2)
    ia2, ib2 = c.front_end()
    ref2 = oracle.candidate_pairs(a, b2)
    _, _, hit2, seg2 = oracle.predicate_pairs(a, b2, ref2)
    hab2, hseg2 = c.hits()
    assert np.array_equal(hab2, ref2[hit2.astype(bool)]) and hseg2.tobytes() == seg2[hit2.astype(bool)].tobytes()
    oa2, _, _ = oracle.classify(b2, oracle.centroids(*a))
    ob2, _, _ = oracle.classify(a, oracle.centroids(*b2))
    assert np.array_equal(ia2, oa2) and np.array_equal(ib2, ob2)
    c.close()


@pytest.mark.slow
def test_comm_c2_two_ranks_equals_single_gpu(ctx):
    import torch
    a, b = meshgen.config_c2()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
    db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    hab, seg = x.hits()
    c = sb.Comm(_devices(2))
    c.set_meshes(a, b)
    ia, ib = c.front_end()
    assert (c.num_candidates, c.num_hits) == (x.num_candidates, x.num_hits) == (7754, 1386)
    cab, cseg = c.hits()
    assert np.array_equal(cab, hab) and cseg.tobytes() == seg.tobytes()
    assert np.array_equal(ia, da.cpu().numpy()) and np.array_equal(ib, db.cpu().numpy())
    c.close(); x.close(); ma.close(); mb.close()


def test_comm_bad_arguments():
    lib = sb.load_library()
    import ctypes as C
    h = C.c_void_p()
    assert lib.sb_comm_create(0, None, C.byref(h)) == 1          # SB_ERR_INVALID
    assert lib.sb_comm_create(1, None, None) == 1
    c = sb.Comm([0])
    with pytest.raises(sb.SolidBooleanError):
        c.front_end()                                              # no meshes yet
    c.close()
