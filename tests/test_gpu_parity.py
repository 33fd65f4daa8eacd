"""GPU (B200): the CUDA path, called through the C ABI, against the oracle and
the committed reference fixtures.  Bit-exact for pairs, codes, flags; segments
are compared bit-for-bit too (north_star allows 1e-12 relative, we get 0)."""
import os
import numpy as np
import pytest

import solidboolean_b200 as sb
from conftest import CASES, load_case, load_synthetic, synthetic_specs
from solidboolean_b200 import meshgen

pytestmark = pytest.mark.gpu

SEG_RTOL = 1e-12  # north_star tolerance for segment endpoints


@pytest.fixture(scope="module")
def ctx():
    c = sb.Context(0)
    yield c
    c.close()


def assert_segments(seg, ref):
    assert seg.shape == ref.shape
    if seg.tobytes() == ref.tobytes():
        return
    scale = np.maximum(np.abs(ref), 1e-300)
    assert np.all(np.abs(seg - ref) <= SEG_RTOL * scale), "segment endpoints differ beyond 1e-12 relative"
    raise AssertionError("segments within tolerance but not bit-identical (expected bit-exact)")


def check_pair(ctx, a, b, out, oracle=None):
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    ab, code = x.candidates()
    assert np.array_equal(ab, out["pairs"]), "candidate-pair set differs"
    assert np.array_equal(code & 1, out["ret"].astype(np.uint8))
    assert np.array_equal((code >> 1) & 1, out["coplanar"].astype(np.uint8))
    hab, seg = x.hits()
    assert np.array_equal(hab, out["pairs"][out["hit"].astype(bool)])
    assert_segments(seg, out["seg_hits"])
    fa, fb = x.face_flags()
    ea = np.zeros(len(a[1]), np.uint8)
    eb = np.zeros(len(b[1]), np.uint8)
    ea[hab[:, 0]] = 1
    eb[hab[:, 1]] = 1
    assert np.array_equal(fa, ea) and np.array_equal(fb, eb)
    ia, pa = ma.classify_faces_against(mb)
    ib, pb = mb.classify_faces_against(ma)
    assert np.array_equal(pa, out["per_axis_a"]) and np.array_equal(ia, out["inside_a"])
    assert np.array_equal(pb, out["per_axis_b"]) and np.array_equal(ib, out["inside_b"])
    if oracle is not None:
        assert ma.normals().tobytes() == oracle.normals(*a).tobytes()
        assert mb.triangle_boxes().tobytes() == oracle.tri_boxes(*b).tobytes()
    x.close(); ma.close(); mb.close()


def test_predicate_known_answers(ctx, golden_kat):
    ret, cop, seg = ctx.tri_tri_batch(golden_kat["tris"])
    assert np.array_equal(ret, golden_kat["ret"])
    assert np.array_equal(cop, golden_kat["coplanar"])
    assert seg.tobytes() == golden_kat["seg"].tobytes()


def test_predicate_random_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(11)
    tris = np.concatenate([
        rng.uniform(-1, 1, (300000, 18)),
        rng.integers(-2, 3, (300000, 18)).astype(np.float64),
        rng.uniform(-1, 1, (100000, 18)).astype(np.float32).astype(np.float64),
        rng.uniform(-1, 1, (50000, 18)) * 1e-150,
        rng.uniform(-1, 1, (50000, 18)) * 1e150,
    ])
    r0, c0, s0 = oracle.tri_tri_batch(tris)
    r1, c1, s1 = ctx.tri_tri_batch(tris)
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1)
    assert s0.tobytes() == s1.tobytes()


@pytest.mark.parametrize("case", CASES)
def test_bundled_cases(ctx, oracle, golden_cases, case):
    a, b, out = load_case(golden_cases, case)
    check_pair(ctx, a, b, out, oracle)


@pytest.mark.parametrize("name", sorted(synthetic_specs()))
def test_synthetic_fixtures(ctx, oracle, golden_synthetic, name):
    a, b, out = load_synthetic(golden_synthetic, name)
    check_pair(ctx, a, b, out, oracle)


def check_bvh(mesh, boxes):
    """Every cluster reachable exactly once; every child box encloses its subtree."""
    t = mesh.bvh()
    K, M = t["cluster_size"], t["num_clusters"]
    n = mesh.num_triangles
    assert M == (n + K - 1) // K
    leaves = t["leaves"]
    order = mesh.order()
    assert sorted(order.tolist()) == list(range(n))
    assert np.array_equal(leaves["ref"][:n], order.astype(np.int32))
    assert np.all(leaves["ref"][n:] == -1)
    lb = boxes[order]
    assert np.all(leaves["lo"][:n] <= lb[:, :3]) and np.all(leaves["hi"][:n] >= lb[:, 3:])
    # tight: conservative rounding moves a bound by at most one float ulp
    assert np.all(leaves["lo"][:n].astype(np.float64) >= lb[:, :3] - np.abs(lb[:, :3]) * 2.0 ** -22 - 1e-37)
    clo = np.full((M, 3), np.inf)
    chi = np.full((M, 3), -np.inf)
    for c in range(M):
        s = slice(c * K, min((c + 1) * K, n))
        clo[c] = leaves["lo"][s].min(axis=0)
        chi[c] = leaves["hi"][s].max(axis=0)
    if M == 1:
        assert t["root"] == -1
        return
    nodes = t["nodes"]
    seen = np.zeros(M, np.int32)
    visited = np.zeros(M - 1, np.int32)

    def walk(ref):
        # returns (lo, hi) of the subtree; iterative to survive deep trees
        stack = [(ref, False)]
        res = {}
        while stack:
            r, done = stack.pop()
            if r < 0:
                c = ~r
                seen[c] += 1
                res[r] = (clo[c], chi[c])
                continue
            if not done:
                visited[r] += 1
                stack.append((r, True))
                stack.append((int(nodes["ref"][2 * r]), False))
                stack.append((int(nodes["ref"][2 * r + 1]), False))
            else:
                lo = np.full(3, np.inf)
                hi = np.full(3, -np.inf)
                for side in (0, 1):
                    cr = int(nodes["ref"][2 * r + side])
                    slo, shi = res[cr]
                    rec = nodes[2 * r + side]
                    assert np.all(rec["lo"] <= slo) and np.all(rec["hi"] >= shi)
                    assert np.array_equal(rec["lo"], slo.astype(np.float32)) and np.array_equal(rec["hi"], shi.astype(np.float32))
                    lo = np.minimum(lo, slo)
                    hi = np.maximum(hi, shi)
                res[r] = (lo, hi)
        return res[ref]

    walk(int(t["root"]))
    assert np.all(seen == 1), "cluster not reached exactly once"
    assert np.all(visited == 1), "internal node not reached exactly once"


@pytest.mark.parametrize("gen", ["ico3", "torus", "slab", "tiny"])
def test_lbvh_invariants(ctx, oracle, gen):
    mesh = {"ico3": lambda: meshgen.icosphere(3), "torus": lambda: meshgen.torus(40, 24),
            "slab": lambda: meshgen.slab(9), "tiny": lambda: meshgen.icosphere(0)}[gen]()
    m = ctx.mesh(*mesh)
    check_bvh(m, oracle.tri_boxes(*mesh))
    b = m.bounds()
    assert np.array_equal(b[:3], mesh[0].min(axis=0)) and np.array_equal(b[3:], mesh[0].max(axis=0))
    m.close()


def test_duplicate_morton_codes_and_degenerate_extent(ctx, oracle):
    """All triangles in one plane / many identical centroids: index tie-break path."""
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64)
    tri = np.tile(np.array([[0, 1, 2]], np.uint32), (300, 1))
    m = ctx.mesh(xyz, tri)
    check_bvh(m, oracle.tri_boxes(xyz, tri))
    x = m.intersect(m)
    assert x.num_candidates == 300 * 300
    ab, code = x.candidates()
    assert np.array_equal(ab, oracle.candidate_pairs((xyz, tri), (xyz, tri)))
    assert np.all((code >> 1) & 1 == 1)  # identical triangles are coplanar
    assert x.num_hits == 0
    x.close(); m.close()


def test_small_and_ragged_inputs(ctx, oracle):
    for k, (nu, nv) in ((0, (3, 3)), (1, (4, 3)), (2, (5, 7))):
        a = meshgen.icosphere(k)
        b = meshgen.torus(nu, nv, R=0.8, r=0.4, center=(0.1, 0.05, 0.02))
        ma, mb = ctx.mesh(*a), ctx.mesh(*b)
        x = ma.intersect(mb)
        ab, code = x.candidates()
        ref = oracle.candidate_pairs(a, b)
        assert np.array_equal(ab, ref)
        ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
        assert np.array_equal(code, (ret | (cop << 1)).astype(np.uint8))
        hab, hseg = x.hits()
        assert np.array_equal(hab, ref[hit.astype(bool)])
        assert hseg.tobytes() == seg[hit.astype(bool)].tobytes()
        ia, pa = ma.classify_faces_against(mb)
        oi, op, _ = oracle.classify(b, oracle.centroids(*a))
        assert np.array_equal(ia, oi) and np.array_equal(pa, op)
        x.close(); ma.close(); mb.close()


def test_single_triangle_and_disjoint(ctx, oracle):
    one = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64), np.array([[0, 1, 2]], np.uint32))
    sph = meshgen.icosphere(2, radius=0.5, center=(0.2, 0.2, 0.0))
    far = meshgen.icosphere(2, center=(10, 10, 10))
    m1, ms, mf = ctx.mesh(*one), ctx.mesh(*sph), ctx.mesh(*far)
    x = m1.intersect(ms)
    assert np.array_equal(x.candidates()[0], oracle.candidate_pairs(one, sph))
    x.close()
    x = ms.intersect(m1)
    assert np.array_equal(x.candidates()[0], oracle.candidate_pairs(sph, one))
    x.close()
    x = ms.intersect(mf)
    assert x.num_candidates == 0 and x.num_hits == 0
    assert x.candidates()[0].shape == (0, 2) and x.hits()[1].shape == (0, 6)
    fa, fb = x.face_flags()
    assert not fa.any() and not fb.any()
    x.close()
    ins, _ = ms.classify_faces_against(mf)
    assert not ins.any()
    for m in (m1, ms, mf):
        m.close()


def test_invalid_arguments(ctx):
    xyz = np.zeros((3, 3))
    with pytest.raises(sb.SolidBooleanError, match="out of range"):
        ctx.mesh(xyz, np.array([[0, 1, 3]], np.uint32))
    bad = ctx.mesh(xyz, np.array([[0, 1, 2], [0, 7, 2]], np.uint32), build=False)   # split upload/build path:
    with pytest.raises(sb.SolidBooleanError, match="out of range"):                   # the build is asynchronous, the
        bad.build()                                                                   # error surfaces at the latest
        bad.normals()                                                                 # with the first use of the mesh
    bad.close()
    m = ctx.mesh(*meshgen.icosphere(1))
    with pytest.raises(sb.SolidBooleanError, match="multiple of 32"):
        m.intersect(m, begin=5, end=40)
    m.close()


def test_explicit_points_vs_oracle(oracle, monkeypatch):
    monkeypatch.setenv("SB_GRID3_EAGER_BELOW", "0")   # small meshes too start with two ray grids (see below)
    ctx = sb.Context(0)
    rng = np.random.default_rng(5)
    for mesh in (meshgen.icosphere(4), meshgen.torus(64, 32, center=(0.013, 0.007, 0.011)),
                 meshgen.slab(16, 1.0, 0.3, tilt=0.2)):
        pts = np.concatenate([rng.uniform(-1.6, 1.6, (20000, 3)),
                              mesh[0][rng.integers(0, len(mesh[0]), 2000)],          # exactly on vertices
                              oracle.centroids(*mesh)[:2000]])                       # exactly on faces
        m = ctx.mesh(*mesh)
        ins, per = m.classify(pts)
        oi, op, _ = oracle.classify(mesh, pts)
        assert np.array_equal(per, op) and np.array_equal(ins, oi)
        # lazy vote (third ray only where the first two disagree) gives the same majority;
        # the on-vertex / on-face points make the axes disagree, so pass 2 really runs
        lazy, none = m.classify(pts, per_axis=False)
        assert none is None and np.array_equal(lazy, oi)
        assert (op[:, 0] != op[:, 1]).sum() > 0
        rays, _ = ctx.classify_stats()
        assert rays == 2 * len(pts) + int((op[:, 0] != op[:, 1]).sum())
        m.close()
        # a fresh mesh has the ray grids of axes 0 and 1 only: the lazy vote lists the undecided
        # points, the third grid is built on demand and a second launch traces their third ray
        m2 = ctx.mesh(*mesh)
        lazy2, _ = m2.classify(pts, per_axis=False)
        assert np.array_equal(lazy2, oi)
        rays2, cands2 = ctx.classify_stats()
        assert rays2 == rays
        # rebuilt meshes keep the third grid: same answer through the single-launch path
        m2.build()
        lazy3, _ = m2.classify(pts, per_axis=False)
        assert np.array_equal(lazy3, oi) and ctx.classify_stats() == (rays2, cands2)
        m2.close()
    ctx.close()


@pytest.mark.parametrize("layers,pitch", [(24, 0.05), (90, 0.02), (40, 1e-7), (70, 3e-6)])
def test_many_layers_hit_list_overflow_path(ctx, oracle, layers, pitch):
    """Rays with more matches than the per-ray staging area take the warp-cooperative
    path: 24 thin slabs stacked along x = 48 crossings for +x rays (keys held in shared
    memory); 90 slabs = 180 crossings (more than 128 distinct keys: global scratch,
    including the host's exact-size retry).  40 slabs 1e-7 apart = 80 matches per ray but two distinct
    PositionKeys, 70 slabs 3e-6 apart = runs of equal keys: the balanced kernel's several-chunk rays
    (sb_classify2.cu) with their short key list."""
    parts_v, parts_t = [], []
    off = 0
    for i in range(layers):
        v, t = meshgen.slab(2, 1.0, 0.01, center=(0, 0, 0))
        rot = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], np.float64)  # slab normal along x
        v = v @ rot.T + np.array([pitch * i, 0, 0])
        parts_v.append(v); parts_t.append(t + off); off += len(v)
    mesh = (np.concatenate(parts_v), np.concatenate(parts_t).astype(np.uint32))
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.uniform(-0.4, 0.4, (500, 3)) * [0, 1, 1] + [-1.0, 0, 0],
                          rng.uniform(-0.6, 1.5, (500, 3)), rng.uniform(-0.6, 1.5, (100000 if layers > 50 else 0, 3))])
    m = ctx.mesh(*mesh)
    ins, per = m.classify(pts)
    oi, op, _ = oracle.classify(mesh, pts)
    assert np.array_equal(per, op) and np.array_equal(ins, oi)
    m.close()


def test_extreme_query_points(ctx, oracle):
    """Non-finite and huge query coordinates (NaN, +-inf, +-DBL_MAX, 1e300, denormals, -0.0): the
    ray-box shortcut's general branch and the quantiser's clamps against the reference's arithmetic."""
    mesh = meshgen.icosphere(3)
    m = ctx.mesh(*mesh)
    big = np.finfo(np.float64).max
    specials = [np.nan, np.inf, -np.inf, big, -big, 1e300, -1e300, 5e-324, -5e-324, -0.0, 0.0, 0.5, -0.5, 2.0]
    pts = np.array([(x, y, z) for x in specials for y in specials for z in (0.0, 0.25, np.nan, -np.inf, big)], np.float64)
    with np.errstate(all="ignore"):
        oi, op, ncand = oracle.classify(mesh, pts)
    ins, per = m.classify(pts)
    assert np.array_equal(per, op) and np.array_equal(ins, oi)
    assert ctx.classify_stats()[1] == ncand
    lazy, _ = m.classify(pts, per_axis=False)
    assert np.array_equal(lazy, oi)
    m.close()


def test_rays_on_cell_borders(ctx, oracle):
    """Points whose DBL_EPSILON-wide ray box straddles a border of the target's ray-grid cells
    (the box then spans two cells): the general warp path, 'first cell only' rule included."""
    mesh = meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    m = ctx.mesh(*mesh)
    gi = m.grid_info()
    lo, hi = mesh[0].min(0), mesh[0].max(0)
    ext = hi - lo
    scl = 32767.999 / ext                      # the device's quantiser (sb_gridq.cuh), same IEEE operations
    rng = np.random.default_rng(77)
    pts = []
    straddling = 0
    for d, cells in ((1, gi["nu"][0]), (2, gi["nv"][0]), (0, gi["nu"][1])):   # y and z borders (x rays), x borders (y rays)
        shift = 15 - int(np.log2(cells))
        for k in rng.integers(1, cells, 300):
            xb = lo[d] + (int(k) << shift) / scl[d]
            for j in range(-4, 5):
                x = xb
                for _ in range(abs(j)):
                    x = np.nextafter(x, np.inf if j > 0 else -np.inf)
                p = rng.uniform(lo - 0.1, hi + 0.1)
                p[d] = x
                q0 = int(np.floor((x - lo[d]) * scl[d])); q1 = int(np.floor(((x + 2.220446049250313e-16) - lo[d]) * scl[d]))
                straddling += (q0 >> shift) != (q1 >> shift)
                pts.append(p)
    pts = np.array(pts)
    assert straddling > 50
    ins, per = m.classify(pts)
    oi, op, ncand = oracle.classify(mesh, pts)
    assert np.array_equal(per, op) and np.array_equal(ins, oi)
    assert ctx.classify_stats()[1] == ncand
    m.close()


def test_general_ray_path_forced(oracle, monkeypatch):
    """SB_CLASSIFY_POOL_LIMIT=2 sends every ray with more than two quantised matches through
    the reference-by-reference warp path (normally only rays spanning several cells or
    overflowing the staging pool): same flags, per-axis bits and candidate counts."""
    monkeypatch.setenv("SB_CLASSIFY_POOL_LIMIT", "2")
    c2 = sb.Context(0)
    rng = np.random.default_rng(31)
    for mesh in (meshgen.torus(64, 32, center=(0.013, 0.007, 0.011)), meshgen.icosphere(4)):
        pts = np.concatenate([rng.uniform(-1.5, 1.5, (20000, 3)), oracle.centroids(*mesh)[:3000]])
        m = c2.mesh(*mesh)
        ins, per = m.classify(pts)
        oi, op, ncand = oracle.classify(mesh, pts)
        assert np.array_equal(per, op) and np.array_equal(ins, oi)
        assert c2.classify_stats()[1] == ncand
        lazy, _ = m.classify(pts, per_axis=False)
        assert np.array_equal(lazy, oi)
        m.close()
    c2.close()


def test_mixed_scale_mesh_big_triangle_list(ctx, oracle):
    """A fine sphere inside a 12-triangle box: the box faces cover the whole ray
    grid (per-axis 'big' list) while the sphere fills ordinary cells."""
    sph = meshgen.icosphere(5, radius=0.4, center=(0.1, 0.05, -0.02))
    s = 1.0
    bx = np.array([[-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s], [-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]], np.float64)
    bt = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [1, 2, 6], [1, 6, 5],
                   [2, 3, 7], [2, 7, 6], [3, 0, 4], [3, 4, 7]], np.uint32)
    mesh = (np.concatenate([sph[0], bx]), np.concatenate([sph[1], bt + len(sph[0])]).astype(np.uint32))
    rng = np.random.default_rng(12)
    pts = np.concatenate([rng.uniform(-1.3, 1.3, (30000, 3)), oracle.centroids(*mesh)[::7]])
    m = ctx.mesh(*mesh)
    ins, per = m.classify(pts)
    oi, op, _ = oracle.classify(mesh, pts)
    assert np.array_equal(per, op) and np.array_equal(ins, oi)
    other = meshgen.icosphere(4, radius=0.7, center=(0.5, 0.4, 0.3))
    mo = ctx.mesh(*other)
    ia, pa = mo.classify_faces_against(m)
    oa, opa, _ = oracle.classify(mesh, oracle.centroids(*other))
    assert np.array_equal(pa, opa) and np.array_equal(ia, oa)
    x = m.intersect(mo)
    assert np.array_equal(x.candidates()[0], oracle.candidate_pairs(mesh, other))
    x.close(); m.close(); mo.close()


def test_flat_and_degenerate_targets(ctx, oracle):
    """Zero-extent dimensions (all triangles in one plane) and a single triangle."""
    rng = np.random.default_rng(2)
    g = np.linspace(-1, 1, 9)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    xyz = np.stack([gx, gy, np.zeros_like(gx)], -1).reshape(-1, 3)
    i = np.arange(8)[:, None]; j = np.arange(8)[None, :]
    a = (i * 9 + j).ravel(); b = ((i + 1) * 9 + j).ravel(); c = ((i + 1) * 9 + j + 1).ravel(); d = (i * 9 + j + 1).ravel()
    tri = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    one = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64), np.array([[0, 1, 2]], np.uint32))
    for mesh in ((xyz, tri), one):
        pts = np.concatenate([rng.uniform(-1.5, 1.5, (5000, 3)), rng.uniform(-1.5, 1.5, (5000, 3)) * [1, 1, 0],
                              rng.uniform(-1.5, 1.5, (2000, 3)) * [1, 0, 1] + [0, 0.25, -2]])
        m = ctx.mesh(*mesh)
        ins, per = m.classify(pts)
        oi, op, _ = oracle.classify(mesh, pts)
        assert np.array_equal(per, op) and np.array_equal(ins, oi)
        m.close()


def test_sharded_ranges_cover_whole(ctx):
    """A's Morton range split in R shards == the unsharded result."""
    a, b = meshgen.icosphere(5), meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    whole = ma.intersect(mb)
    wab, wcode = whole.candidates()
    whab, wseg = whole.hits()
    n = ma.num_triangles
    for R in (2, 3, 8):
        cuts = [((n * r // R) // 32) * 32 for r in range(R)] + [n]
        abs_, codes, habs, segs = [], [], [], []
        for r in range(R):
            x = ma.intersect(mb, begin=cuts[r], end=cuts[r + 1])
            ab, code = x.candidates(); hab, seg = x.hits()
            abs_.append(ab); codes.append(code); habs.append(hab); segs.append(seg)
            x.close()
        ab = np.concatenate(abs_); code = np.concatenate(codes)
        o = np.lexsort((ab[:, 1], ab[:, 0]))
        assert np.array_equal(ab[o], wab) and np.array_equal(code[o], wcode)
        hab = np.concatenate(habs); seg = np.concatenate(segs)
        o = np.lexsort((hab[:, 1], hab[:, 0]))
        assert np.array_equal(hab[o], whab) and seg[o].tobytes() == wseg.tobytes()
    whole.close(); ma.close(); mb.close()


def test_front_end_overlapped_equals_separate_calls(ctx):
    """sb_front_end(_range) = sb_intersect + both sb_classify_faces, stages overlapped."""
    import torch
    a, b = meshgen.icosphere(5), meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    ref_x = ma.intersect(mb)
    ia, _ = ma.classify_faces_against(mb)
    ib, _ = mb.classify_faces_against(ma)
    da = torch.full((len(a[1]),), 7, dtype=torch.uint8, device="cuda")
    db = torch.full((len(b[1]),), 7, dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    assert np.array_equal(da.cpu().numpy(), ia) and np.array_equal(db.cpu().numpy(), ib)
    for u, v in zip(x.candidates() + x.hits(), ref_x.candidates() + ref_x.hits()):
        assert u.tobytes() == v.tobytes()
    rays, cands = ctx.classify_stats()
    q = len(a[1]) + len(b[1])
    assert 2 * q <= rays <= 3 * q and cands > 0   # lazy vote: third ray only where two disagree
    # sharded: two halves of each query set fill disjoint parts of the flag arrays
    da.fill_(0); db.fill_(0)
    na, nb = len(a[1]), len(b[1])
    ca, cb = (na // 2) // 32 * 32, (nb // 2) // 32 * 32
    parts = []
    for ra, rb in (((0, ca), (0, cb)), ((ca, na), (cb, nb))):
        ta = torch.zeros_like(da); tb = torch.zeros_like(db)
        xs = sb.Isect.front_end(ma, mb, ta.data_ptr(), tb.data_ptr(), a_range=ra, b_range=rb)
        da += ta; db += tb
        parts.append(xs.hits()[0]); xs.close()
    assert np.array_equal(da.cpu().numpy(), ia) and np.array_equal(db.cpu().numpy(), ib)
    hab = np.concatenate(parts)
    assert np.array_equal(hab[np.lexsort((hab[:, 1], hab[:, 0]))], ref_x.hits()[0])
    x.close(); ref_x.close(); ma.close(); mb.close()


def test_predicate_exit_histogram_and_fp64_probe(ctx, oracle):
    a, b = meshgen.icosphere(4), meshgen.icosphere(4, center=(0.71, 0.13, 0.07))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    paths = x.path_counts()
    ab, code = x.candidates()
    assert sum(paths) == x.num_candidates
    assert paths[4] == int(((code & 1) == 1).sum() - ((code >> 1) & 1 & (code & 1)).sum())  # segments = ret && !coplanar
    assert paths[2] == int(((code >> 1) & 1).sum())                                        # coplanar flag set
    assert paths[0] + paths[1] + paths[3] == int((code == 0).sum())
    nofma, fma = ctx.fp64_peak()
    assert 1e3 < nofma < 1e5 and fma > 1.5 * nofma     # B200: tens of TFLOP/s, FMA = 2 flops per issue
    x.close(); ma.close(); mb.close()


def test_no_sort_flag_same_set(ctx):
    a, b = meshgen.icosphere(4), meshgen.icosphere(4, center=(0.71, 0.13, 0.07))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    s = ma.intersect(mb)
    u = ma.intersect(mb, flags=sb.ISECT_NO_SORT)
    sab, scode = s.candidates(); uab, ucode = u.candidates()
    o = np.lexsort((uab[:, 1], uab[:, 0]))
    assert np.array_equal(uab[o], sab) and np.array_equal(ucode[o], scode)
    shab, sseg = s.hits(); uhab, useg = u.hits()
    o = np.lexsort((uhab[:, 1], uhab[:, 0]))
    assert np.array_equal(uhab[o], shab) and useg[o].tobytes() == sseg.tobytes()
    s.close(); u.close(); ma.close(); mb.close()


def test_rebuild_is_idempotent(ctx):
    a, b = meshgen.icosphere(4), meshgen.torus(48, 24, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x0 = ma.intersect(mb); r0 = (x0.candidates(), x0.hits()); x0.close()
    o0 = ma.order()
    ma.build(); mb.build()
    assert np.array_equal(ma.order(), o0)
    x1 = ma.intersect(mb); r1 = (x1.candidates(), x1.hits()); x1.close()
    for u, v in zip(r0[0] + r0[1], r1[0] + r1[1]):
        assert u.tobytes() == v.tobytes()
    ma.close(); mb.close()


def test_config_c2_vs_oracle(ctx, oracle):
    """BASELINE config 2: two offset icospheres, 81,920 + 81,920 triangles."""
    a, b = meshgen.config_c2()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x = ma.intersect(mb)
    ab, code = x.candidates()
    ref = oracle.candidate_pairs(a, b)
    assert np.array_equal(ab, ref)
    ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
    assert np.array_equal(code, (ret | (cop << 1)).astype(np.uint8))
    hab, hseg = x.hits()
    assert np.array_equal(hab, ref[hit.astype(bool)]) and hseg.tobytes() == seg[hit.astype(bool)].tobytes()
    assert (x.num_candidates, x.num_hits) == (7754, 1386)  # BASELINE.md probe table
    ia, pa = ma.classify_faces_against(mb)
    oi, op, _ = oracle.classify(b, oracle.centroids(*a))
    assert np.array_equal(pa, op) and np.array_equal(ia, oi)
    x.close(); ma.close(); mb.close()


@pytest.mark.slow
def test_config_c4_dense_pairs_properties(ctx, oracle):
    """BASELINE config 4 proxy: near-coincident icospheres k=7, 327,680 x2 triangles,
    2.26 M candidate pairs -- the predicate-heavy case.  Full candidate / hit / segment equality
    with the oracle, sampled classification."""
    import torch
    a, b = meshgen.config_c4()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
    db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    ab, code = x.candidates()
    hab, seg = x.hits()
    ref = oracle.candidate_pairs(a, b)
    assert len(ref) > 2_000_000 and np.array_equal(ab, ref)
    ret, cop, hit, oseg = oracle.predicate_pairs(a, b, ab)
    assert np.array_equal(code, (ret | (cop << 1)).astype(np.uint8))
    assert np.array_equal(hab, ab[hit.astype(bool)]) and seg.tobytes() == oseg[hit.astype(bool)].tobytes()
    assert int(sum(x.path_counts())) == len(ab)      # every pair left the predicate through one of its five exits
    ia, ib = da.cpu().numpy(), db.cpu().numpy()
    rng = np.random.default_rng(14)
    sa = rng.choice(len(a[1]), 20000, replace=False)
    sb_ = rng.choice(len(b[1]), 20000, replace=False)
    oa, _, _ = oracle.classify(b, oracle.centroids(*a)[sa])
    ob, _, _ = oracle.classify(a, oracle.centroids(*b)[sb_])
    assert np.array_equal(ia[sa], oa) and np.array_equal(ib[sb_], ob)
    x.close(); ma.close(); mb.close()


@pytest.mark.slow
def test_config_c3_full_size_properties(ctx, oracle):
    """BASELINE config 3 (1,310,720 + 1,048,576 triangles): size-independent properties
    plus sampled oracle checks (the full oracle classification would take minutes)."""
    import torch
    a, b = meshgen.config_c3()
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
    db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    assert (x.num_candidates, x.num_hits) == (36125, 9606)          # BASELINE.md probe table
    ab, code = x.candidates()
    hab, seg = x.hits()
    # sortedness, uniqueness, hits are exactly the candidates with ret=1, coplanar=0
    key = ab[:, 0].astype(np.int64) << 32 | ab[:, 1]
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(hab, ab[code == 1])
    # the full candidate set against the oracle's own accelerator (exact equality)
    assert np.array_equal(ab, oracle.candidate_pairs(a, b))
    ret, cop, hit, oseg = oracle.predicate_pairs(a, b, ab)
    assert np.array_equal(code, (ret | (cop << 1)).astype(np.uint8))
    assert seg.tobytes() == oseg[hit.astype(bool)].tobytes()
    # symmetry: B against A finds the transposed candidate set
    y = mb.intersect(ma)
    ba, _ = y.candidates()
    t = ba[:, ::-1]
    assert np.array_equal(t[np.lexsort((t[:, 1], t[:, 0]))], ab)
    y.close()
    # classification: totals, a 20K-face sample per mesh against the oracle, and
    # every face far inside / far outside by construction
    ia, ib = da.cpu().numpy(), db.cpu().numpy()
    assert (int(ia.sum()), int(ib.sum())) == (454612, 465529)
    rng = np.random.default_rng(8)
    sa = rng.choice(len(a[1]), 20000, replace=False)
    sb_ = rng.choice(len(b[1]), 20000, replace=False)
    oa, _, _ = oracle.classify(b, oracle.centroids(*a)[sa])
    ob, _, _ = oracle.classify(a, oracle.centroids(*b)[sb_])
    assert np.array_equal(ia[sa], oa) and np.array_equal(ib[sb_], ob)
    cb = oracle.centroids(*b)
    r = np.linalg.norm(cb, axis=1)
    assert np.all(ib[r < 0.99] == 1) and np.all(ib[r > 1.01] == 0)   # torus faces inside / outside the unit sphere
    # the lazy vote equals the full three-axis vote
    full, per = ma.classify_faces_against(mb)
    assert np.array_equal(full, ia) and np.array_equal((per.sum(axis=1) >= 2).astype(np.uint8), ia)
    x.close(); ma.close(); mb.close()


def _fullsize_pins(name):
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")) as f:
        return json.load(f)[name]


def _hash(a):
    from oracle import fnv1a64
    return "%016x" % fnv1a64(np.ascontiguousarray(a).tobytes())


def _check_against_fullsize_pins(ctx, name, a, b):
    """EVERY output of the front end at full size against the unmodified reference
    (tests/golden/fullsize.json, written by tests/golden/make_fullsize.py from oracle/_ref over all
    faces): sorted candidate pairs, per-pair (ret, coplanar), hit pairs, segments bit for bit,
    per-face flags and per-axis bits of all faces of both meshes."""
    import torch
    pin = _fullsize_pins(name)
    assert (len(a[1]), len(b[1])) == (pin["tris_a"], pin["tris_b"])
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda")
    db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    assert (x.num_candidates, x.num_hits) == (pin["n_pairs"], pin["n_hits"])
    ab, code = x.candidates()
    hab, seg = x.hits()
    assert _hash(ab.astype("<u8")) == pin["pairs"]
    assert _hash(code.astype(np.uint8)) == pin["codes"]
    assert _hash(hab.astype("<u8")) == pin["hits"]
    assert _hash(seg.astype("<f8")) == pin["seg"]
    ia, ib = da.cpu().numpy(), db.cpu().numpy()
    assert (int(ia.sum()), int(ib.sum())) == (pin["inside_a_count"], pin["inside_b_count"])
    assert _hash(ia) == pin["inside_a"] and _hash(ib) == pin["inside_b"]       # lazy vote, all faces
    fa, pa = ma.classify_faces_against(mb)                                          # all three rays, all faces
    fb, pb = mb.classify_faces_against(ma)
    assert _hash(pa.astype(np.uint8)) == pin["per_axis_a"] and _hash(pb.astype(np.uint8)) == pin["per_axis_b"]
    assert np.array_equal(fa, ia) and np.array_equal(fb, ib)
    x.close(); ma.close(); mb.close()


@pytest.mark.slow
def test_config_c3_every_output_vs_reference_pins(ctx):
    _check_against_fullsize_pins(ctx, "c3", *meshgen.config_c3())


@pytest.mark.slow
def test_config_c4_stated_size_vs_reference_pins(ctx):
    """BASELINE config 4 at its stated size: near-coincident icospheres k=8, 1,310,720 x2 triangles,
    8,238,598 candidate pairs."""
    _check_against_fullsize_pins(ctx, "c4k8", *meshgen.config_c4(k=8))


def test_mesh_update_new_geometry_into_sized_lists(ctx, oracle):
    """sb_mesh_update + sb_mesh_build: new coordinates of the same counts into meshes whose ray-grid reference
    lists were sized for the old geometry.  The rebuild fills within the old capacity and reports its counts;
    the first use checks them (more references than the capacity -> the lists are sized again).  Every frame
    must equal the oracle on that frame's geometry -- including frames that need MORE references than the
    first one (stretched so that the triangle boxes straddle many more cells) and big-list triangles that
    appear only later."""
    a0, b0 = meshgen.icosphere(4), meshgen.torus(64, 32, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a0), ctx.mesh(*b0)
    rng = np.random.default_rng(5)

    def frames():
        yield a0[0] * np.array([1.0, 1.0, 0.05]), b0[0]                       # squashed: anisotropic boxes
        yield a0[0] + rng.normal(0, 0.02, a0[0].shape), b0[0] * 1.1          # noisy: larger, overlapping boxes
        va = a0[0].copy(); va[::97] *= 40.0                                   # a few far vertices: huge triangles (big lists)
        yield va, b0[0]
        yield a0[0], b0[0]                                                    # and back
    for va, vb in frames():
        xa = np.ascontiguousarray(va, np.float64); xb = np.ascontiguousarray(vb, np.float64)
        ma.update(xa.ctypes.data, 0); mb.update(xb.ctypes.data, 0)
        ma.build(); mb.build()
        a, b = (xa, a0[1]), (xb, b0[1])
        x = ma.intersect(mb)
        ab, code = x.candidates()
        ref = oracle.candidate_pairs(a, b)
        assert np.array_equal(ab, ref)
        ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
        hab, hseg = x.hits()
        assert np.array_equal(hab, ref[hit.astype(bool)]) and hseg.tobytes() == seg[hit.astype(bool)].tobytes()
        ia, pa = ma.classify_faces_against(mb)
        ib, pb = mb.classify_faces_against(ma)
        oa, opa, _ = oracle.classify(b, oracle.centroids(*a))
        ob, opb, _ = oracle.classify(a, oracle.centroids(*b))
        assert np.array_equal(pa, opa) and np.array_equal(ia, oa)
        assert np.array_equal(pb, opb) and np.array_equal(ib, ob)
        x.close()
    ma.close(); mb.close()


def test_vertex_indices_far_apart_use_the_index_fallback(ctx, oracle):
    """The classifier's triangle record packs the three vertex indices as (i0, i1 - i0, i2 - i0) with 21-bit
    differences; a triangle whose corners lie more than 2^20 vertices apart keeps the all-ones word and the kernel
    reads `tri`.  Target with 2.2 M vertices, every triangle using one far copy of a corner."""
    v, t = meshgen.icosphere(3)
    far = 2_200_000
    big = np.repeat(v[:1], far + len(v), axis=0)
    big[:len(v)] = v
    big[far:far + len(v)] = v
    t2 = t.copy()
    t2[:, 1] += far                                    # second corner from the far copy: |i1 - i0| > 2^20
    tgt = (np.ascontiguousarray(big), np.ascontiguousarray(t2.astype(np.uint32)))
    q = meshgen.torus(48, 24, center=(0.013, 0.007, 0.011))
    mt, mq = ctx.mesh(*tgt), ctx.mesh(*q)
    iq, pq = mq.classify_faces_against(mt)
    oi, op, _ = oracle.classify(tgt, oracle.centroids(*q))
    assert np.array_equal(pq, op) and np.array_equal(iq, oi)
    x = mq.intersect(mt)
    ab, _ = x.candidates()
    assert np.array_equal(ab, oracle.candidate_pairs(q, tgt))
    x.close(); mt.close(); mq.close()


def _check_hit_edges(a, b, hab, seg, tags):
    """Every tagged end point lies on the edge the tag names (distance to the edge's segment below 1e-9 of the
    triangle's size): independent, geometric check of sb_isect_hit_edges."""
    assert len(tags) == len(hab) and np.all(tags & 0x80)
    va, ta = a
    vb, tb = b
    for shift, pts in ((0, seg[:, 0:3]), (4, seg[:, 3:6])):
        edge = (tags >> shift) & 3
        on_b = ((tags >> (shift + 2)) & 1).astype(bool)
        assert np.all(edge < 3)
        tri = np.where(on_b[:, None], tb[hab[:, 1]], ta[hab[:, 0]])          # vertex ids of the owning triangle
        verts = np.where(on_b[:, None, None], vb[tb[hab[:, 1]]], va[ta[hab[:, 0]]])
        i0 = edge
        i1 = (edge + 1) % 3
        p0 = verts[np.arange(len(tags)), i0]
        p1 = verts[np.arange(len(tags)), i1]
        d = p1 - p0
        t = np.einsum("ij,ij->i", pts - p0, d) / np.maximum(np.einsum("ij,ij->i", d, d), 1e-300)
        tc = np.clip(t, 0.0, 1.0)
        dist = np.linalg.norm(pts - (p0 + tc[:, None] * d), axis=1)
        size = np.linalg.norm(verts.max(axis=1) - verts.min(axis=1), axis=1)
        assert np.all(dist <= 1e-9 * size), (float((dist / size).max()), int(np.argmax(dist / size)))
        assert tri.shape[1] == 3


def test_hit_edge_tags_name_the_edges_the_end_points_lie_on(ctx):
    """the edge ids carried out of the predicate, on inputs that exercise every permutation:
    two soups of random triangles (thousands of hits in general position), C2's curve, a mesh pair with shared planes."""
    rng = np.random.default_rng(21)

    def soup(n, scale):
        c = rng.uniform(-1, 1, (n, 1, 3))
        v = (c + rng.normal(0, scale, (n, 3, 3))).reshape(-1, 3)
        return np.ascontiguousarray(v), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    cases = [(soup(3000, 0.15), soup(3000, 0.15)), meshgen.config_c2(),
             (meshgen.icosphere(3), meshgen.torus(48, 24, center=(0.013, 0.007, 0.011)))]
    seen = set()
    for a, b in cases:
        ma, mb = ctx.mesh(*a), ctx.mesh(*b)
        x = ma.intersect(mb)
        hab, seg = x.hits()
        tags = x.hit_edges()
        assert len(hab) > 100
        _check_hit_edges(a, b, hab.astype(np.int64), seg, tags)
        seen |= set(int(t) & 0x77 for t in tags)
        x.close(); ma.close(); mb.close()
    # all four branches of CONSTRUCT_INTERSECTION occur: (T1,T2), (T2,T2), (T1,T1), (T2,T1) as (source, target) owners
    owners = set(((t >> 2) & 1, (t >> 6) & 1) for t in seen)
    assert owners == {(0, 1), (1, 1), (0, 0), (1, 0)}
    assert len(seen) >= 20                                                   # and most (edge, edge) combinations


def test_parked_mesh_is_revived_with_new_geometry_and_topology(oracle):
    """sb_mesh_destroy parks plain meshes; sb_mesh_upload of the same counts hands the object out again (arena, streams,
    captured rebuild, reference-list sizes verified against the new geometry).  A caller that makes new meshes of the same
    sizes per operation -- different coordinates AND a different triangle order each time -- must get each operation's own
    result."""
    c = sb.Context(0)
    rng = np.random.default_rng(11)
    a0, b0 = meshgen.icosphere(4), meshgen.torus(64, 32, center=(0.013, 0.007, 0.011))
    for it in range(4):
        perm_a, perm_b = rng.permutation(len(a0[1])), rng.permutation(len(b0[1]))
        rot = np.array([[np.cos(0.3 * it), -np.sin(0.3 * it), 0], [np.sin(0.3 * it), np.cos(0.3 * it), 0], [0, 0, 1.0]])
        a = (np.ascontiguousarray(a0[0] @ rot.T * (1.0 + 0.1 * it)), np.ascontiguousarray(a0[1][perm_a]))
        b = (np.ascontiguousarray(b0[0] + 0.02 * it), np.ascontiguousarray(b0[1][perm_b]))
        ma, mb = c.mesh(*a), c.mesh(*b)
        x = ma.intersect(mb)
        ab, _ = x.candidates()
        ref = oracle.candidate_pairs(a, b)
        assert np.array_equal(ab, ref)
        _, _, hit, seg = oracle.predicate_pairs(a, b, ref)
        hab, hseg = x.hits()
        assert np.array_equal(hab, ref[hit.astype(bool)]) and hseg.tobytes() == seg[hit.astype(bool)].tobytes()
        ia, pa = ma.classify_faces_against(mb)
        oa, opa, _ = oracle.classify(b, oracle.centroids(*a))
        assert np.array_equal(pa, opa) and np.array_equal(ia, oa)
        assert np.array_equal(ma.normals(), oracle.normals(*a))
        x.close(); ma.close(); mb.close()       # parked; the next round's meshes of the same counts revive them
    c.close()


@pytest.mark.slow
@pytest.mark.parametrize("last", ["b", "a"])
@pytest.mark.parametrize("pool_limit", [None, "8"])
def test_streamed_classification_of_a_just_uploaded_mesh(oracle, monkeypatch, last, pool_limit):
    """sb_mesh_upload / sb_mesh_update send the index triples in chunks; the first sb_front_end after it classifies
    the faces of the mesh that arrived LAST in their original order, chunk by chunk, straight from the uploaded
    arrays (not from that mesh's build).  Flags must equal the oracle's and the ordinary (Morton-order) path's --
    with either mesh last, through the general kernel too (SB_CLASSIFY_POOL_LIMIT), with an OPEN target whose
    first two rays disagree (undecided list + third grid on demand) and after sb_mesh_update of new coordinates."""
    import torch
    monkeypatch.setenv("SB_GRID3_EAGER_BELOW", "0")
    monkeypatch.setenv("SB_STREAM_CLASSIFY", "1")   # (off by default: no gain at C3, the GPU is busy either way)
    if pool_limit:
        monkeypatch.setenv("SB_CLASSIFY_POOL_LIMIT", pool_limit)
    ctx = sb.Context(0)
    a = meshgen.icosphere(7)                                            # 327,680 faces: four upload chunks
    b = meshgen.torus(256, 128, center=(0.013, 0.007, 0.011))          # 65,536 faces: two
    # an open target: a band of the sphere's faces removed -> x and y rays of some points disagree
    cen = oracle.centroids(*a)
    a_open = (a[0], np.ascontiguousarray(a[1][np.abs(cen[:, 2] - 0.31) > 0.05]))
    for A, B in ((a, b), (a_open, b)):
        keep = [np.ascontiguousarray(A[0], np.float64), np.ascontiguousarray(A[1], np.uint32),
                np.ascontiguousarray(B[0], np.float64), np.ascontiguousarray(B[1], np.uint32)]
        mk = lambda i: sb.Mesh.from_pointers(ctx, keep[2 * i].ctypes.data, len(keep[2 * i]), keep[2 * i + 1].ctypes.data,
                                             len(keep[2 * i + 1]), build=False, keep=keep)
        if last == "b":
            ma = mk(0); mb = mk(1)
        else:
            mb = mk(1); ma = mk(0)
        ma.build(); mb.build()
        da = torch.full((len(A[1]),), 7, dtype=torch.uint8, device="cuda")
        db = torch.full((len(B[1]),), 7, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        oa, _, _ = oracle.classify(B, oracle.centroids(*A))
        ob, _, _ = oracle.classify(A, oracle.centroids(*B))
        assert np.array_equal(da.cpu().numpy(), oa) and np.array_equal(db.cpu().numpy(), ob)
        rays1 = ctx.classify_stats()
        ref_hits = x.hits()
        x.close()
        # the same meshes again, nothing uploaded in between: the ordinary path (Morton order of both builds)
        da.fill_(7); db.fill_(7); torch.cuda.synchronize()   # (torch's stream is not ordered with the library's)
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        assert np.array_equal(da.cpu().numpy(), oa) and np.array_equal(db.cpu().numpy(), ob)
        assert ctx.classify_stats() == rays1 and all(u.tobytes() == v.tobytes() for u, v in zip(x.hits(), ref_hits))
        x.close()
        # new coordinates into the same meshes (sb_mesh_update), both orders of arrival
        va = np.ascontiguousarray(A[0] * np.array([1.0, 0.9, 1.1])); vb = np.ascontiguousarray(B[0] * 1.05)
        for m, v, t in ((ma, va, keep[1]), (mb, vb, keep[3])) if last == "b" else ((mb, vb, keep[3]), (ma, va, keep[1])):
            m.update(v.ctypes.data, t.ctypes.data)
        ma.build(); mb.build()
        da.fill_(7); db.fill_(7); torch.cuda.synchronize()
        x = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
        oa, _, _ = oracle.classify((vb, B[1]), oracle.centroids(va, A[1]))
        ob, _, _ = oracle.classify((va, A[1]), oracle.centroids(vb, B[1]))
        assert np.array_equal(da.cpu().numpy(), oa) and np.array_equal(db.cpu().numpy(), ob)
        x.close(); ma.close(); mb.close()
    ctx.close()


def test_front_end_with_host_outputs(ctx, oracle):
    """sb_front_end_host = sb_front_end + the copies to the host, each enqueued behind the stream that produced it:
    same flags, same hit list; a landing buffer that is too small for the hit list is left alone."""
    import torch
    a, b = meshgen.icosphere(5), meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    da = torch.zeros(len(a[1]), dtype=torch.uint8, device="cuda"); db = torch.zeros(len(b[1]), dtype=torch.uint8, device="cuda")
    ref = sb.Isect.front_end(ma, mb, da.data_ptr(), db.data_ptr())
    rab, rseg = ref.hits()
    assert len(rab) > 0
    for pinned in (True, False):
        ha = torch.full((len(a[1]),), 9, dtype=torch.uint8); hb = torch.full((len(b[1]),), 9, dtype=torch.uint8)
        cap = len(rab) + 5
        hab = torch.full((2 * cap,), -1, dtype=torch.int32); hseg = torch.full((6 * cap,), -1.0, dtype=torch.float64)
        if pinned:
            ha, hb, hab, hseg = ha.pin_memory(), hb.pin_memory(), hab.pin_memory(), hseg.pin_memory()
        x = sb.Isect.front_end_host(ma, mb, ha.data_ptr(), hb.data_ptr(), hab.data_ptr(), hseg.data_ptr(), cap)
        assert (x.num_candidates, x.num_hits) == (ref.num_candidates, ref.num_hits)
        assert np.array_equal(ha.numpy(), da.cpu().numpy()) and np.array_equal(hb.numpy(), db.cpu().numpy())
        n = x.num_hits
        assert np.array_equal(hab.numpy()[:2 * n].reshape(-1, 2).astype(np.uint32), rab)
        assert hseg.numpy()[:6 * n].tobytes() == rseg.tobytes() and int(hab[2 * n]) == -1
        x.close()
        # too small a landing buffer: untouched, sb_isect_hits still delivers
        hab.fill_(-1)
        x = sb.Isect.front_end_host(ma, mb, ha.data_ptr(), hb.data_ptr(), hab.data_ptr(), hseg.data_ptr(), len(rab) - 1)
        assert int(hab[0]) == -1 and np.array_equal(x.hits()[0], rab)
        x.close()
        # flags only
        ha.fill_(9)
        x = sb.Isect.front_end_host(ma, mb, ha.data_ptr(), hb.data_ptr())
        assert np.array_equal(ha.numpy(), da.cpu().numpy())
        x.close()
    oa, _, _ = oracle.classify(b, oracle.centroids(*a))
    assert np.array_equal(da.cpu().numpy(), oa)
    ref.close(); ma.close(); mb.close()


@pytest.mark.parametrize("optimistic", ["1", "0"])
def test_front_end_right_after_update_is_enqueued_before_the_counts_are_known(oracle, monkeypatch, optimistic):
    """sb_mesh_update + sb_mesh_build + sb_front_end_host back to back: the front end is enqueued behind the rebuilds
    without waiting for their reference counts (SB_OPTIMISTIC, default on) and checks them afterwards.  Frames whose
    references still fit, a frame that needs MORE references than the lists hold and one where big-list triangles
    appear (both: the run is repeated after a proper rebuild) -- every frame must equal the oracle."""
    import torch
    monkeypatch.setenv("SB_OPTIMISTIC", optimistic)
    ctx = sb.Context(0)
    a0, b0 = meshgen.icosphere(5), meshgen.torus(96, 48, center=(0.013, 0.007, 0.011))
    ma, mb = ctx.mesh(*a0), ctx.mesh(*b0)
    rng = np.random.default_rng(11)
    ha = torch.zeros(len(a0[1]), dtype=torch.uint8).pin_memory(); hb = torch.zeros(len(b0[1]), dtype=torch.uint8).pin_memory()
    cap = 1 << 16
    hab = torch.zeros(2 * cap, dtype=torch.int32).pin_memory(); hseg = torch.zeros(6 * cap, dtype=torch.float64).pin_memory()

    def frames():
        yield a0[0] * 1.01, b0[0]                                              # fits
        yield a0[0] * np.array([1.0, 1.0, 0.05]), b0[0]                       # squashed: anisotropic boxes
        yield a0[0] + rng.normal(0, 0.02, a0[0].shape), b0[0] * 1.1          # noisy: many more references than the lists hold
        va = a0[0].copy(); va[::97] *= 40.0                                   # a few far vertices: huge triangles (big lists appear)
        yield va, b0[0]
        yield a0[0], b0[0]                                                    # and back (big lists vanish)
        yield a0[0] * 0.99, b0[0] * 1.01                                      # fits again
    for va, vb in frames():
        xa = np.ascontiguousarray(va, np.float64); xb = np.ascontiguousarray(vb, np.float64)
        ma.update(xa.ctypes.data, 0); mb.update(xb.ctypes.data, 0)
        ma.build(); mb.build()
        ha.fill_(9); hb.fill_(9)
        x = sb.Isect.front_end_host(ma, mb, ha.data_ptr(), hb.data_ptr(), hab.data_ptr(), hseg.data_ptr(), cap)
        a, b = (xa, a0[1]), (xb, b0[1])
        ref = oracle.candidate_pairs(a, b)
        ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
        n = x.num_hits
        assert x.num_candidates == len(ref) and n == int(hit.sum()) and n <= cap
        assert np.array_equal(hab.numpy()[:2 * n].reshape(-1, 2).astype(np.uint32), ref[hit.astype(bool)])
        assert hseg.numpy()[:6 * n].tobytes() == seg[hit.astype(bool)].tobytes()
        oa, _, _ = oracle.classify(b, oracle.centroids(*a))
        ob, _, _ = oracle.classify(a, oracle.centroids(*b))
        assert np.array_equal(ha.numpy(), oa) and np.array_equal(hb.numpy(), ob)
        x.close()
    ma.close(); mb.close(); ctx.close()


def test_build_head_runs_beside_the_upload(oracle, monkeypatch, golden_cases):
    """sb_mesh_upload / sb_mesh_update from host memory start the head of the build while the arrays arrive (bounds and
    padded vertices behind the coordinates, normals / Morton keys / digit counts behind each chunk of index triples);
    sb_mesh_build goes on from the sort.  With the size threshold lowered every mesh here takes that path: results must
    equal the oracle's, for new meshes, for updated ones (new coordinates AND new triangle order), and the index check
    of the per-triangle kernel must still be reported."""
    monkeypatch.setenv("SB_EARLY_PREP", "1")        # (off by default: 1 % at C3)
    monkeypatch.setenv("SB_EARLY_PREP_MIN", "1")
    ctx = sb.Context(0)
    rng = np.random.default_rng(3)
    for a, b in ((meshgen.icosphere(4), meshgen.torus(64, 32, center=(0.013, 0.007, 0.011))),
                 (meshgen.icosphere(5), meshgen.icosphere(5, center=(0.21, 0.13, -0.05)))):
        ma, mb = ctx.mesh(*a), ctx.mesh(*b)
        for frame in range(3):
            if frame:
                pa, pb = rng.permutation(len(a[1])), rng.permutation(len(b[1]))
                a = (np.ascontiguousarray(a[0] * (1.0 + 0.03 * frame)), np.ascontiguousarray(a[1][pa]))
                b = (np.ascontiguousarray(b[0] + 0.01 * frame), np.ascontiguousarray(b[1][pb]))
                ma.update(a[0].ctypes.data, a[1].ctypes.data); mb.update(b[0].ctypes.data, b[1].ctypes.data)
                ma.build(); mb.build()
            x = ma.intersect(mb)
            ref = oracle.candidate_pairs(a, b)
            ret, cop, hit, seg = oracle.predicate_pairs(a, b, ref)
            assert np.array_equal(x.candidates()[0], ref)
            hab, hseg = x.hits()
            assert np.array_equal(hab, ref[hit.astype(bool)]) and hseg.tobytes() == seg[hit.astype(bool)].tobytes()
            ia, pa_ = ma.classify_faces_against(mb)
            oa, opa, _ = oracle.classify(b, oracle.centroids(*a))
            assert np.array_equal(pa_, opa) and np.array_equal(ia, oa)
            assert ma.normals().tobytes() == oracle.normals(*a).tobytes()
            x.close()
        ma.close(); mb.close()
    bad = meshgen.icosphere(4)
    tri = bad[1].copy(); tri[100, 1] = len(bad[0]) + 5
    with pytest.raises(sb.SolidBooleanError):
        ctx.mesh(bad[0], tri)
    ctx.close()
