"""CPU: pin the C oracle (oracle/sb_oracle.c) to the reference.

Two anchors: (1) the committed golden fixtures produced by the unmodified
reference (tests/golden/make_golden.py), incl. the SURVEY section-4 hashes;
(2) when oracle/_ref/libsbref.so is present, the reference itself, live.
"""
import numpy as np
import pytest

from conftest import CASES, load_case, load_synthetic, synthetic_specs
from oracle import Ref, fnv1a64

SURVEY_HASHES = {  # SURVEY.md section 4 (pairs, hits, segments)
    "simple-ring": ("417a7631554d1725", "f5c8fb94fe697d25", "28894104c51956e7"),
    "cube-sphere": ("281f0cf346df2da5", "178e3433f46c031a", "13756fda0f219a33"),
    "complex": ("d51f9f0a7f34aea1", "b78cfc299e750cf1", "10a237b628700c83"),
    "addax-and-meerkat": ("3d338a306d83169f", "c8c2a4ba6b4338ad", "d0a81ab21cb2d3fc"),
}
SURVEY_COUNTS = {  # tris A, tris B, P, H, insideA, insideB
    "simple-ring": (12, 12, 16, 8, 0, 6),
    "cube-sphere": (12, 48, 40, 20, 0, 13),
    "complex": (100, 3688, 5008, 274, 1, 1606),
    "addax-and-meerkat": (3458, 2952, 2026, 208, 20, 318),
}


def _hash_pairs(p):
    return "%016x" % fnv1a64(np.asarray(p, np.uint64).astype("<u8").tobytes())


def check_against(oracle, a, b, out):
    pairs = oracle.candidate_pairs(a, b)
    assert np.array_equal(pairs, out["pairs"])
    ret, cop, hit, seg = oracle.predicate_pairs(a, b, pairs)
    assert np.array_equal(ret, out["ret"])
    assert np.array_equal(cop, out["coplanar"])
    assert np.array_equal(hit, out["hit"])
    assert seg[hit.astype(bool)].tobytes() == out["seg_hits"].tobytes()
    ia, pa, _ = oracle.classify(b, oracle.centroids(*a))
    ib, pb, _ = oracle.classify(a, oracle.centroids(*b))
    assert np.array_equal(ia, out["inside_a"]) and np.array_equal(pa, out["per_axis_a"])
    assert np.array_equal(ib, out["inside_b"]) and np.array_equal(pb, out["per_axis_b"])
    return pairs, hit, seg


@pytest.mark.parametrize("case", CASES)
def test_bundled_cases_match_reference_fixtures(oracle, golden_cases, golden_json, case):
    a, b, out = load_case(golden_cases, case)
    pairs, hit, seg = check_against(oracle, a, b, out)
    na, nb, P, H, ina, inb = SURVEY_COUNTS[case]
    assert (len(a[1]), len(b[1]), len(pairs), int(hit.sum())) == (na, nb, P, H)
    assert (int(out["inside_a"].sum()), int(out["inside_b"].sum())) == (ina, inb)
    hp, hh, hs = SURVEY_HASHES[case]
    assert _hash_pairs(pairs) == hp
    assert _hash_pairs(pairs[hit.astype(bool)]) == hh
    assert "%016x" % fnv1a64(seg[hit.astype(bool)].tobytes()) == hs
    assert golden_json["cases"][case]["hash_pairs"] == hp


@pytest.mark.parametrize("name", sorted(synthetic_specs()))
def test_synthetic_match_reference_fixtures(oracle, golden_synthetic, golden_json, name):
    a, b, out = load_synthetic(golden_synthetic, name)
    pairs, hit, _ = check_against(oracle, a, b, out)
    meta = golden_json["synthetic"][name]
    assert (len(pairs), int(hit.sum())) == (meta["P"], meta["H"])
    assert _hash_pairs(pairs) == meta["hash_pairs"]


def test_predicate_known_answers(oracle, golden_kat, golden_json):
    ret, cop, seg = oracle.tri_tri_batch(golden_kat["tris"])
    assert np.array_equal(ret, golden_kat["ret"])
    assert np.array_equal(cop, golden_kat["coplanar"])
    assert seg.tobytes() == golden_kat["seg"].tobytes()
    assert "%016x" % fnv1a64(seg.tobytes()) == golden_json["tritri_kat"]["hash_seg"]
    # the crafted set must reach the coplanar branch and both outcomes
    assert cop.sum() > 1000 and 0 < ret.sum() < len(ret)


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_vs_live_reference_random(oracle):
    R = Ref.get()
    rng = np.random.default_rng(7)
    tris = np.concatenate([
        rng.uniform(-1, 1, (20000, 18)),
        rng.integers(-2, 3, (20000, 18)).astype(np.float64),
        rng.uniform(-1, 1, (5000, 18)).astype(np.float32).astype(np.float64),
    ])
    r0, c0, s0 = R.tri_tri_batch(tris)
    r1, c1, s1 = oracle.tri_tri_batch(tris)
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and s0.tobytes() == s1.tobytes()


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_vs_live_reference_mesh(oracle):
    from solidboolean_b200 import meshgen
    R = Ref.get()
    a = meshgen.icosphere(4, center=(0.1, 0.0, 0.0))
    b = meshgen.torus(40, 20, R=0.9, r=0.3, center=(0.013, 0.007, 0.011))
    ma, mb = R.mesh(*a), R.mesh(*b)
    op = R.op(ma, mb)
    pr = op.search()
    prs = pr[np.lexsort((pr[:, 1], pr[:, 0]))]
    assert np.array_equal(oracle.candidate_pairs(a, b), prs)
    ret, cop, hit, seg, _ = op.predicate(prs)
    oret, ocop, ohit, oseg = oracle.predicate_pairs(a, b, prs)
    assert np.array_equal(ret, oret) and np.array_equal(cop, ocop) and np.array_equal(hit, ohit)
    assert seg.tobytes() == oseg.tobytes()
    assert oracle.normals(*a).tobytes() == ma.normals().tobytes()
    assert oracle.tri_boxes(*b).tobytes() == mb.boxes().tobytes()
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1.5, 1.5, (3000, 3))
    i0, p0, _ = op.classify(1, pts)
    i1, p1, _ = oracle.classify(b, pts)
    assert np.array_equal(i0, i1) and np.array_equal(p0, p1)


# ---- uncut triangles + half-edge map ---------------------------------

def _h(a, dt):
    return "%016x" % fnv1a64(np.ascontiguousarray(a).astype(dt).tobytes())


def oracle_uncut_pair(oracle, a, b, ca, cb):
    """Both meshes as combine() chains them (reference src/solidboolean.cpp:411-421): the second
    mesh's vertices sit behind the first's, its triangles behind the first's uncut ones."""
    ra = oracle.uncut_half_edges(a[1], ca, 0, 0)
    rb = oracle.uncut_half_edges(b[1], cb, len(a[0]), len(ra["face"]))
    return ra, rb


@pytest.fixture(scope="module")
def uncut_golden():
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "uncut.json")) as f:
        return json.load(f)


def test_uncut_half_edges_match_reference_fixtures(oracle, uncut_golden):
    from conftest import uncut_inputs
    inputs = uncut_inputs()
    assert sorted(inputs) == sorted(uncut_golden)
    for name, (a, b, ca, cb) in inputs.items():
        ra, rb = oracle_uncut_pair(oracle, a, b, ca, cb)
        tris = np.concatenate([a[1][ra["face"]], b[1][rb["face"]] + len(a[0])]).astype(np.uint32)
        g = uncut_golden[name]
        assert _h(tris, "<u4") == g["triangles_hash"], name
        for r, gg in ((ra, g["a"]), (rb, g["b"])):
            assert (r["ok"], len(r["face"]), len(r["keys"])) == (gg["ok"], gg["n_triangles"], gg["n_half_edges"]), name
            assert _h(r["keys"], "<u8") == gg["keys_hash"], name
            assert _h(r["owner"], "<u4") == gg["owner_hash"], name
            assert _h(r["adj"], "<i4") == gg["adj_hash"], name
            assert int((r["adj"] < 0).sum()) == gg["boundary_edges"], name
            assert (r["components"], _h(r["label"], "<u4")) == (gg["groups"], gg["label_hash"]), name


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built")
def test_uncut_half_edges_match_live_reference(oracle):
    from solidboolean_b200 import meshgen
    rng = np.random.default_rng(5)
    a = meshgen.icosphere(4)
    b = meshgen.torus(40, 20, center=(0.2, 0.0, 0.1))
    R = Ref.get()
    op = R.op(R.mesh(*a), R.mesh(*b))
    for trial in range(3):
        ca = (rng.random(len(a[1])) < 0.1 * trial).astype(np.uint8)
        cb = (rng.random(len(b[1])) < 0.3).astype(np.uint8)
        (ra, rb), tris = op.uncut(ca, cb)
        oa, ob = oracle_uncut_pair(oracle, a, b, ca, cb)
        for r, o in ((ra, oa), (rb, ob)):
            assert r["ok"] and o["ok"]
            assert np.array_equal(r["keys"], o["keys"]) and np.array_equal(r["owner"], o["owner"])
        assert np.array_equal(tris, np.concatenate([a[1][oa["face"]], b[1][ob["face"]] + len(a[0])]))
        t = a[1][oa["face"]].astype(np.uint64)
        ft = np.stack([np.stack([t[:, (k + 1) % 3], t[:, k]], -1) for k in range(3)], 1).reshape(-1, 2)
        assert np.array_equal(op.uncut_lookup(0, ft).reshape(-1, 3), oa["adj"])
        for w, o in ((0, oa), (1, ob)):
            label, groups = op.uncut_groups(w, len(o["face"]))
            assert groups == o["components"] and np.array_equal(label, o["label"])


# ---- per-triangle intersection contexts ------------------------------

def _context_inputs(golden_cases, golden_synthetic):
    out = {c: load_case(golden_cases, c) for c in CASES}
    out.update({n: load_synthetic(golden_synthetic, n) for n in sorted(synthetic_specs())})
    return out


def test_cut_contexts_match_reference_fixtures(oracle, golden_cases, golden_synthetic):
    """tests/golden/contexts.npz = what the unmodified reference's own pair loop built, observed at
    ReTriangulator::setEdges (oracle/ref_hook.cpp), pair list ascending."""
    import os
    from conftest import GOLDEN, check_contexts_against
    fx = np.load(os.path.join(GOLDEN, "contexts.npz"))
    total = 0
    for name, (a, b, out) in _context_inputs(golden_cases, golden_synthetic).items():
        hits = out["pairs"][out["hit"].astype(bool)]
        key = name.replace("-", "_")
        for w, s in enumerate("ab"):
            o = oracle.cut_contexts(hits, out["seg_hits"], w)
            assert np.array_equal(o["tri"], np.unique(hits[:, w]))
            check_contexts_against(o, lambda k: fx["%s__%s_%s" % (key, s, k)], name + ":" + s)
            total += len(fx["%s__%s_tri" % (key, s)])
    assert total > 1500      # 1,558 contexts of the reference in the fixture


@pytest.mark.skipif(not (Ref.available() and __import__("os").path.exists(__import__("oracle").HOOK_SO)),
                    reason="oracle/_ref (reference + hook) not built")
def test_cut_contexts_match_live_reference(oracle):
    from conftest import check_contexts_against
    from oracle.ref_contexts import capture
    from solidboolean_b200 import meshgen
    a = meshgen.icosphere(3, round_to_float=True)
    b = meshgen.torus(24, 12, center=(0.31, 0.02, 0.05))
    cap = capture(a, b, True)
    assert cap["ok"][0] == 1 and len(cap["hits"]) > 50
    assert np.all(np.diff(cap["hits"][:, 0].astype(np.int64) << 32 | cap["hits"][:, 1]) > 0)   # the loop saw ascending pairs
    for w, s in enumerate("ab"):
        o = oracle.cut_contexts(cap["hits"], cap["seg"], w)
        assert len(cap[s + "_tri"]) == len(o["tri"])
        check_contexts_against(o, lambda k: cap[s + "_" + k], s)
    # the reference's own (tree-traversal) order gives the same context SETS for this input,
    # but not necessarily the same point numbering: only the sorted run is the contract
    cap2 = capture(a, b, False)
    assert sorted(map(tuple, cap2["hits"].tolist())) == list(map(tuple, cap["hits"].tolist()))


# ---- buildFaceGroups with loops (the part still on the host) ------------------

def _boundary_loops(tri, sel):
    """Closed vertex cycles along the mesh edges that separate the selected triangles from the rest."""
    he = {}
    for t in np.flatnonzero(sel):
        for k in range(3):
            he[(int(tri[t, k]), int(tri[t, (k + 1) % 3]))] = t
    nxt = {u: v for (u, v) in he if (v, u) not in he}
    loops, seen = [], set()
    for s in sorted(nxt):
        if s in seen:
            continue
        loop, c = [], s
        while c not in seen:
            seen.add(c)
            loop.append(c)
            c = nxt[c]
        loops.append(loop)
    return loops


def _group_cases():
    from solidboolean_b200 import meshgen
    for name, (v, t) in {"ico3": meshgen.icosphere(3), "torus": meshgen.torus(24, 12)}.items():
        cen = v[t].mean(axis=1)
        for sel_name, sel in {"z>0": cen[:, 2] > 0.0, "x>0.3": cen[:, 0] > 0.3, "caps": np.abs(cen[:, 2]) > 0.2}.items():
            yield name + ":" + sel_name, t, sel, _boundary_loops(t, sel)


def test_face_groups_with_loops_split_where_the_loops_run(oracle):
    for name, t, sel, loops in _group_cases():
        r = oracle.uncut_half_edges(t, None, 0, 0)
        group, order, n = oracle.face_groups(t, r["keys"], r["owner"], loops, 0, len(t))
        assert n == 2 * len(loops) and len(order) == len(t) and np.all(group != 0xffffffff), name
        # loop k runs with the selected triangles on its left: they land in the even groups
        assert np.all((group[sel] % 2) == 0) and np.all((group[~sel] % 2) == 1), name
        # with a third of the faces missing from the map the flood stops there: leftovers open new groups
        cut = (np.arange(len(t)) % 3 == 0).astype(np.uint8)
        rc = oracle.uncut_half_edges(t, cut, 0, 0)
        tri_new = t[rc["face"]]
        g2, o2, n2 = oracle.face_groups(tri_new, rc["keys"], rc["owner"], [], 0, len(tri_new))
        assert n2 == rc["components"] and len(o2) == len(tri_new), name
        first = {}
        for tt in o2:
            first.setdefault(int(g2[tt]), int(tt))
        assert np.array_equal(np.array([first[int(g)] for g in g2], np.uint32), rc["label"]), name


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built")
def test_face_groups_with_loops_match_live_reference(oracle):
    R = Ref.get()
    for name, t, sel, loops in _group_cases():
        r = oracle.uncut_half_edges(t, None, 0, 0)
        for start, count in ((0, len(t)), (0, 0), (len(t) // 2, len(t) // 3)):
            og, oo, on = oracle.face_groups(t, r["keys"], r["owner"], loops, start, count)
            rg, ro, rn = R.face_groups(t, r["keys"], r["owner"], loops, start, count)
            assert on == rn and np.array_equal(og, rg), (name, start, count)
            # the reference's groups, concatenated = the oracle's assignment order, stably sorted by group
            assert np.array_equal(oo[np.argsort(og[oo], kind="stable")], ro), (name, start, count)
        # a map with holes (cut faces) and no loops: leftover groups only
        cut = (np.arange(len(t)) % 3 == 0).astype(np.uint8)
        rc = oracle.uncut_half_edges(t, cut, 0, 0)
        tri_new = t[rc["face"]]
        og, oo, on = oracle.face_groups(tri_new, rc["keys"], rc["owner"], [], 0, len(tri_new))
        rg, ro, rn = R.face_groups(tri_new, rc["keys"], rc["owner"], [], 0, len(tri_new))
        assert on == rn and np.array_equal(og, rg) and np.array_equal(oo[np.argsort(og[oo], kind="stable")], ro), name
