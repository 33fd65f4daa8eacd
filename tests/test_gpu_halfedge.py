"""GPU (B200): uncut triangles + half-edge map (sb_isect_uncut / sb_mesh_uncut) against the
oracle restatement of SolidBoolean::addUnintersectedTriangles (reference
src/solidboolean.cpp:250-286) and the committed reference fixtures (tests/golden/uncut.json).
Index work: everything bit-exact."""
import json
import os

import numpy as np
import pytest

import solidboolean_b200 as sb
from conftest import GOLDEN, uncut_inputs
from oracle import fnv1a64
from solidboolean_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def _h(a, dt):
    return "%016x" % fnv1a64(np.ascontiguousarray(a).astype(dt).tobytes())


def check_uncut(u, ref, tri, vertex_offset):
    """u: sb.Uncut; ref: oracle.uncut_half_edges(...) of the same inputs."""
    assert (u.ok, u.num_triangles, u.num_half_edges) == (ref["ok"], len(ref["face"]), len(ref["keys"]))
    face, tri3 = u.triangles()
    assert np.array_equal(face, ref["face"])
    assert np.array_equal(tri3, np.asarray(tri, np.uint32)[ref["face"]] + np.uint32(vertex_offset))
    keys, owner = u.half_edges()
    assert np.array_equal(keys, ref["keys"]) and np.array_equal(owner, ref["owner"])
    assert np.array_equal(u.adjacency(), ref["adj"])
    label, groups = u.components()
    assert groups == ref["components"] and np.array_equal(label, ref["label"])
    return keys, owner


def test_uncut_fixtures_vs_reference_and_oracle(ctx, oracle):
    with open(os.path.join(GOLDEN, "uncut.json")) as f:
        golden = json.load(f)
    for name, (a, b, ca, cb) in uncut_inputs().items():
        ma, mb = ctx.mesh(*a, build=False), ctx.mesh(*b, build=False)   # the geometry is all this stage reads
        ua = ma.uncut(ca, 0, 0)
        ub = mb.uncut(cb, len(a[0]), ua.num_triangles)
        ra = oracle.uncut_half_edges(a[1], ca, 0, 0)
        rb = oracle.uncut_half_edges(b[1], cb, len(a[0]), len(ra["face"]))
        g = golden[name]
        for u, r, m, voff, gg in ((ua, ra, a, 0, g["a"]), (ub, rb, b, len(a[0]), g["b"])):
            keys, owner = check_uncut(u, r, m[1], voff)
            # ... and straight against the numbers the unmodified reference produced
            assert (u.ok, u.num_triangles, u.num_half_edges) == (gg["ok"], gg["n_triangles"], gg["n_half_edges"]), name
            assert (_h(keys, "<u8"), _h(owner, "<u4")) == (gg["keys_hash"], gg["owner_hash"]), name
            assert _h(u.adjacency(), "<i4") == gg["adj_hash"], name
            label, groups = u.components()
            assert (groups, _h(label, "<u4")) == (gg["groups"], gg["label_hash"]), name
        ua.close(); ub.close(); ma.close(); mb.close()


def test_uncut_from_intersection_flags(ctx, oracle):
    """sb_isect_uncut takes the cut faces from the intersection's own device flags
    (m_firstIntersectedFaces / m_secondIntersectedFaces, reference src/solidboolean.cpp:318-319)."""
    a, b = meshgen.config_c2()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    fa, fb = x.face_flags()
    hab, _ = x.hits()
    assert int(fa.sum()) == len(np.unique(hab[:, 0])) and int(fb.sum()) == len(np.unique(hab[:, 1]))
    ua = x.uncut(0, 0, 0)
    ub = x.uncut(1, len(a[0]), ua.num_triangles)
    ra = oracle.uncut_half_edges(a[1], fa, 0, 0)
    rb = oracle.uncut_half_edges(b[1], fb, len(a[0]), len(ra["face"]))
    check_uncut(ua, ra, a[1], 0)
    check_uncut(ub, rb, b[1], len(a[0]))
    assert ua.num_triangles == len(a[1]) - int(fa.sum())
    # closed manifold input: the only missing neighbours are the cut faces
    adj = ua.adjacency()
    t = a[1]
    cut_neighbours = 0
    he = {}
    for f in np.flatnonzero(fa):
        for k in range(3):
            he[(int(t[f, k]), int(t[f, (k + 1) % 3]))] = f
    face, _ = ua.triangles()
    for j in np.flatnonzero((adj < 0).any(axis=1)):
        for k in range(3):
            if adj[j, k] < 0:
                assert (int(t[face[j], (k + 1) % 3]), int(t[face[j], k])) in he
                cut_neighbours += 1
    assert cut_neighbours == int((adj < 0).sum())
    ua.close(); ub.close(); x.close(); ma.close(); mb.close()


def test_uncut_edge_cases(ctx, oracle):
    # empty mesh, one triangle, ragged sizes around the 2048-face tile and the 8-face word
    tor = meshgen.torus(64, 33)
    rng = np.random.default_rng(11)
    for n in (0, 1, 7, 8, 9, 2047, 2048, 2049, 4100, len(tor[1])):
        tri = tor[1][:n]
        m = ctx.mesh(tor[0], tri, build=False)
        for cut in (None, (rng.random(n) < 0.5).astype(np.uint8), np.ones(n, np.uint8)):
            u = m.uncut(cut, 5, 17)
            check_uncut(u, oracle.uncut_half_edges(tri, cut, 5, 17), tri, 5)
            u.close()
        m.close()
    # vertex offsets close to the 32-bit limit (33 key bits per half: an 8-pass sort)
    m = ctx.mesh(*tor, build=False)
    off = (1 << 32) - len(tor[0])
    u = m.uncut(None, off, 0)
    keys, owner = u.half_edges()
    r = oracle.uncut_half_edges(tor[1], None, off, 0)
    assert np.array_equal(keys, r["keys"]) and np.array_equal(owner, r["owner"]) and np.array_equal(u.adjacency(), r["adj"])
    assert np.array_equal(u.components()[0], r["label"])
    u.close()
    with pytest.raises(sb.SolidBooleanError):
        m.uncut(None, off + 1, 0)
    m.close()


@pytest.mark.slow
def test_uncut_config_c3_full_size(ctx, oracle):
    """BASELINE config 3 (1,310,720 + 1,048,576 triangles): the whole map against the oracle,
    plus the properties a closed manifold must show."""
    a, b = meshgen.config_c3()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    fa, fb = x.face_flags()
    ua = x.uncut(0, 0, 0)
    ub = x.uncut(1, len(a[0]), ua.num_triangles)
    ra = oracle.uncut_half_edges(a[1], fa, 0, 0)
    rb = oracle.uncut_half_edges(b[1], fb, len(a[0]), len(ra["face"]))
    for u, r, m, voff in ((ua, ra, a, 0), (ub, rb, b, len(a[0]))):
        keys, owner = check_uncut(u, r, m[1], voff)
        assert u.ok and np.all(np.diff(keys.astype(np.uint64)) > 0)          # sorted, no repeated half-edge
        adj = u.adjacency()
        # symmetry: my neighbour across an edge has me as a neighbour
        base = u.half_edges()[1].min() if u.num_half_edges else 0
        j, k = np.nonzero(adj >= 0)
        back = adj[adj[j, k] - base]
        assert np.all((back == (j + base)[:, None]).any(axis=1))
    assert int((ua.adjacency() < 0).sum()) > 0 and ua.num_triangles + int(fa.sum()) == len(a[1])
    ua.close(); ub.close(); x.close(); ma.close(); mb.close()


def test_uncut_rejects_out_of_range_indices(ctx):
    tor = meshgen.torus(16, 8)
    bad = tor[1].copy()
    bad[5, 1] = len(tor[0]) + 3
    m = ctx.mesh(tor[0], bad, build=False)
    with pytest.raises(sb.SolidBooleanError):
        m.uncut(None, 0, 0)
    m.close()


@pytest.mark.slow
def test_uncut_config_c4_many_fragments(ctx, oracle):
    """BASELINE config 4 proxy (near-coincident icospheres, 327,680 x 2): a large share of the
    faces is cut, the uncut rest falls apart into many face groups."""
    a, b = meshgen.config_c4()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    fa, fb = x.face_flags()
    assert int(fa.sum()) > 1000 and int(fb.sum()) > 1000
    ua = x.uncut(0, 0, 0)
    ub = x.uncut(1, len(a[0]), ua.num_triangles)
    ra = oracle.uncut_half_edges(a[1], fa, 0, 0)
    rb = oracle.uncut_half_edges(b[1], fb, len(a[0]), len(ra["face"]))
    check_uncut(ua, ra, a[1], 0)
    check_uncut(ub, rb, b[1], len(a[0]))
    ua.close(); ub.close()
    # ... and with a random third of the faces cut: thousands of fragments, single triangles included,
    # through both node orders of the union-find (Morton order of the built mesh / face order)
    rng = np.random.default_rng(21)
    cut = (rng.random(len(a[1])) < 0.35).astype(np.uint8)
    r = oracle.uncut_half_edges(a[1], cut, 7, 3)
    assert r["components"] > 1000
    u = ma.uncut(cut, 7, 3)
    check_uncut(u, r, a[1], 7)
    u.close()
    raw = ctx.mesh(*a, build=False)
    u = raw.uncut(cut, 7, 3)
    check_uncut(u, r, a[1], 7)
    u.close(); raw.close()
    x.close(); ma.close(); mb.close()


def _host_groups(tris, fences):
    """Checker for sb_uncut_face_groups: connected components of `tris` (node = row) through opposite half-edges that
    are not fences (either direction), label = lowest node of the component.  Plain union-find on the host."""
    n = len(tris)
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    fence = set()
    for a, b in fences:
        fence.add((int(a), int(b))); fence.add((int(b), int(a)))
    owner = {}
    for t, tri in enumerate(tris):
        for k in range(3):
            owner[(int(tri[k]), int(tri[(k + 1) % 3]))] = t
    for t, tri in enumerate(tris):
        for k in range(3):
            a, b = int(tri[k]), int(tri[(k + 1) % 3])
            if (a, b) in fence:
                continue
            o = owner.get((b, a))
            if o is not None and o != t:
                ra, rb = find(t), find(o)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    return np.array([find(t) for t in range(n)], np.uint32)


@pytest.mark.parametrize("voff,toff", [(0, 0), (1000, 77)])
def test_face_groups_over_uncut_triangles_and_pieces(ctx, voff, toff):
    """SURVEY 8f row 3 on the device: a band of faces around the equator of an icosphere plays the cut triangles (their
    own triangles handed back as `pieces`, as a trivial retriangulation would), two closed vertex rings inside the band
    are the intersection loops.  The groups must be the connected components the host checker finds: the two caps
    (uncut components + the band pieces on their side of a ring) and the strip between the rings."""
    v, t = meshgen.icosphere(4)
    m = ctx.mesh(v, t)
    z = v[t].mean(axis=1)[:, 2]
    cut = (np.abs(z) < 0.25).astype(np.uint8)
    u = m.uncut(cut, voff, toff)
    face, tri3 = u.triangles()
    pieces = (t[cut.astype(bool)] + voff).astype(np.uint32)
    # fences: the edges of the band faces that cross z = +-0.1 ... need CLOSED curves made of mesh edges: take every edge
    # whose two adjacent faces lie on different sides of the plane z = c (the boundary of {faces with centroid z > c})
    def ring(c):
        side = z > c
        own = {}
        for f, tri in enumerate(t):
            for k in range(3):
                own[(int(tri[k]), int(tri[(k + 1) % 3]))] = f
        out = []
        for (a, b), f in own.items():
            g = own.get((b, a))
            if g is not None and side[f] and not side[g]:
                out.append((a + voff, b + voff))
        return out
    fences = np.array(ring(0.1) + ring(-0.1), np.uint32)
    assert len(fences) > 20
    lu, lp, n = u.face_groups(pieces, fences)
    nodes = np.concatenate([tri3, pieces])                      # uncut triangles first, then the pieces
    ref = _host_groups(nodes, fences)
    assert np.array_equal(lu, ref[:len(tri3)]) and np.array_equal(lp, ref[len(tri3):])
    assert n == len(np.unique(ref)) == 3
    # no fences: everything is one group; no pieces: the uncut components themselves
    lu1, lp1, n1 = u.face_groups(pieces, np.zeros((0, 2), np.uint32))
    assert n1 == 1 and np.all(lu1 == 0) and np.all(lp1 == 0)
    lu2, _, n2 = u.face_groups(np.zeros((0, 3), np.uint32), np.zeros((0, 2), np.uint32))
    lab, ncomp = u.components()
    assert n2 == ncomp == 2 and np.array_equal(lu2, lab - toff)
    u.close(); m.close()


def test_face_groups_random_cuts_and_fences(ctx):
    """Random third of the faces cut (thousands of uncut fragments), random fences on cut-face edges."""
    rng = np.random.default_rng(3)
    v, t = meshgen.torus(96, 48)
    m = ctx.mesh(v, t)
    cut = (rng.random(len(t)) < 0.33).astype(np.uint8)
    u = m.uncut(cut, 0, 0)
    _, tri3 = u.triangles()
    pieces = t[cut.astype(bool)].astype(np.uint32)
    pick = pieces[rng.random(len(pieces)) < 0.5]
    fences = np.concatenate([pick[:, [0, 1]], pick[::3, [1, 2]]]).astype(np.uint32)
    lu, lp, n = u.face_groups(pieces, fences)
    ref = _host_groups(np.concatenate([tri3, pieces]), fences)
    assert np.array_equal(lu, ref[:len(tri3)]) and np.array_equal(lp, ref[len(tri3):])
    assert n == len(np.unique(ref))
    u.close(); m.close()


def test_face_groups_and_hit_edges_reject_bad_arguments(ctx):
    import ctypes as C
    lib = sb.load_library()
    v, t = meshgen.icosphere(2)
    m = ctx.mesh(v, t)
    u = m.uncut(None, 0, 0)
    n = C.c_size_t(0)
    lu = np.zeros(u.num_triangles, np.uint32)
    # pieces announced but no array / no output array
    assert lib.sb_uncut_face_groups(u.h, None, 4, None, 0, lu.ctypes.data_as(C.c_void_p), None, C.byref(n)) == 1
    pc = np.zeros((2, 3), np.uint32)
    assert lib.sb_uncut_face_groups(u.h, pc.ctypes.data_as(C.c_void_p), 2, None, 0, lu.ctypes.data_as(C.c_void_p), None, C.byref(n)) == 1
    assert lib.sb_uncut_face_groups(None, None, 0, None, 0, None, None, C.byref(n)) == 1
    assert lib.sb_uncut_face_groups(u.h, None, 0, None, 3, lu.ctypes.data_as(C.c_void_p), None, C.byref(n)) == 1   # fences without an array
    assert lib.sb_isect_hit_edges(None, None) == 1
    # degenerate but valid: pieces that touch nothing are groups of their own
    far = np.array([[10_000, 10_001, 10_002]], np.uint32)
    lu2, lp2, n2 = u.face_groups(far, np.zeros((0, 2), np.uint32))
    assert n2 == 2 and lp2[0] == u.num_triangles and np.all(lu2 == 0)
    u.close(); m.close()
