"""GPU (B200): per-triangle intersection contexts (sb_isect_contexts) against the oracle
restatement of the pair-loop body of SolidBoolean::combine (reference src/solidboolean.cpp:296-339)
and against what the unmodified reference itself built (tests/golden/contexts.npz, observed with
oracle/ref_hook.cpp).  Points bit for bit, numbering and relations exactly."""
import os

import numpy as np
import pytest

import solidboolean_b200 as sb
from conftest import CASES, GOLDEN, check_contexts_against, load_case, load_synthetic, synthetic_specs
from solidboolean_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def assert_same(dev, ref, what):
    for k in ("tri", "point_start", "edge_start", "edges"):
        assert np.array_equal(dev[k], ref[k]), (what, k)
    assert dev["points"].tobytes() == ref["points"].tobytes(), (what, "points")


def check_pair(ctx, oracle, a, b, what):
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    hab, seg = x.hits()
    out = []
    for w in (0, 1):
        dev = x.contexts(w)
        assert_same(dev, oracle.cut_contexts(hab, seg, w), "%s side %d" % (what, w))
        assert np.array_equal(dev["tri"], np.unique(hab[:, w]))
        out.append(dev)
    x.close(); ma.close(); mb.close()
    return out


def test_contexts_fixtures_vs_reference_and_oracle(ctx, oracle, golden_cases, golden_synthetic):
    fx = np.load(os.path.join(GOLDEN, "contexts.npz"))
    inputs = {c: load_case(golden_cases, c)[:2] for c in CASES}
    inputs.update({n: load_synthetic(golden_synthetic, n)[:2] for n in sorted(synthetic_specs())})
    for name, (a, b) in inputs.items():
        devs = check_pair(ctx, oracle, a, b, name)
        key = name.replace("-", "_")
        for dev, s in zip(devs, "ab"):   # ... and straight against the reference's own contexts
            check_contexts_against(dev, lambda k: fx["%s__%s_%s" % (key, s, k)], name + ":" + s)


def test_contexts_edge_cases(ctx, oracle):
    # no hits at all; a single hit; many segments through one big triangle (a long run)
    far = meshgen.icosphere(1, center=(5.0, 0.0, 0.0))
    near = meshgen.icosphere(1)
    ma, mb = ctx.mesh(*near), ctx.mesh(*far)
    x = ma.intersect(mb)
    for w in (0, 1):
        d = x.contexts(w)
        assert len(d["tri"]) == 0 and list(d["point_start"]) == [0] and list(d["edge_start"]) == [0]
    x.close(); ma.close(); mb.close()
    big = (np.array([[-3.0, -3.0, 0.0], [3.0, -3.0, 0.0], [0.0, 3.0, 0.0], [0.0, 0.0, -4.0]]),
           np.array([[0, 1, 2], [0, 3, 1], [1, 3, 2], [2, 3, 0]], np.uint32))        # a tetrahedron with one large face
    check_pair(ctx, oracle, big, meshgen.icosphere(4, center=(0.0, 0.0, 0.05)), "big face x sphere")
    check_pair(ctx, oracle, meshgen.icosphere(4, center=(0.0, 0.0, 0.05)), big, "sphere x big face")


def test_contexts_config_c2(ctx, oracle):
    check_pair(ctx, oracle, *meshgen.config_c2(), "c2")


@pytest.mark.slow
def test_contexts_config_c3_and_c4(ctx, oracle):
    devs = check_pair(ctx, oracle, *meshgen.config_c3(), "c3")
    assert sum(len(d["tri"]) for d in devs) > 5000
    check_pair(ctx, oracle, *meshgen.config_c4(), "c4")


def test_contexts_and_uncut_reject_bad_arguments(ctx):
    import ctypes as C
    a, b = meshgen.icosphere(2), meshgen.icosphere(2, center=(0.4, 0.1, 0.0))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    lib, h = x.lib, sb._vp()
    assert lib.sb_isect_contexts(None, 0, C.byref(h)) == 1          # SB_ERR_INVALID
    assert lib.sb_isect_contexts(x.h, 2, C.byref(h)) == 1
    assert lib.sb_isect_contexts(x.h, 0, None) == 1
    assert lib.sb_isect_uncut(x.h, 5, 0, 0, C.byref(h)) == 1
    assert lib.sb_mesh_uncut(None, None, 0, 0, C.byref(h)) == 1
    assert lib.sb_uncut_counts(None, None, None, None) == 1 and lib.sb_cuts_counts(None, None, None, None) == 1
    assert b"null" in lib.sb_last_error() or b"cuts" in lib.sb_last_error()
    # hits left in emission order: "first seen" would not be defined
    y = ma.intersect(mb, flags=1)                                    # SB_ISECT_NO_SORT
    assert lib.sb_isect_contexts(y.h, 0, C.byref(h)) == 1
    assert b"ascending" in lib.sb_last_error()
    y.close(); x.close(); ma.close(); mb.close()
