"""Batch meshes (sb_batch_upload, BASELINE configs[4] / SURVEY 8e "C5"): many small jobs through ONE
build and ONE front end must give, job by job, exactly what the reference gives for that job alone
(reference test/main.cpp:92-103 runs one SolidBoolean per call).  Checked against the CPU oracle:
candidate pairs, predicate codes, hit pairs, segments (bit for bit) and per-face inside flags."""
import numpy as np
import pytest

import solidboolean_b200 as sb
from solidboolean_b200 import meshgen

pytestmark = pytest.mark.gpu


def ragged_jobs():
    """Jobs of different shapes and sizes that all overlap in space, one far away, one without
    triangles on one side, one pair that does not touch."""
    jobs = [meshgen.config_c5_job(j, k=3) for j in range(12)]
    jobs.append((meshgen.icosphere(2), meshgen.torus(24, 12, center=(0.013, 0.007, 0.011))))
    jobs.append((meshgen.torus(32, 16, center=(0.1, 0.0, 0.05)), meshgen.icosphere(3, radius=0.9)))
    jobs.append((meshgen.slab(6, 1.0, 0.3, tilt=0.2), meshgen.icosphere(2, center=(0.2, 0.1, 0.0))))
    jobs.append((meshgen.icosphere(1, center=(40.0, -35.0, 20.0)), meshgen.icosphere(2, center=(40.3, -35.0, 20.1))))
    jobs.append((meshgen.icosphere(2), (np.zeros((3, 3)), np.zeros((0, 3), np.uint32))))      # nothing to meet
    jobs.append((meshgen.icosphere(2), meshgen.icosphere(2, center=(5.0, 0.0, 0.0))))         # disjoint
    jobs += [meshgen.config_c5_job(100 + j, k=2, round_to_float=False) for j in range(20)]
    return jobs


def check_batch(ctx, oracle, jobs, sample=None):
    import torch
    pitch = sb.batch_pitch(*[m[0] for ab in jobs for m in ab])
    A = sb.Mesh.batch(ctx, [a for a, _ in jobs], pitch)
    B = sb.Mesh.batch(ctx, [b for _, b in jobs], pitch)
    nA, nB = A.num_triangles, B.num_triangles
    da = torch.zeros(max(nA, 1), dtype=torch.uint8, device="cuda")
    db = torch.zeros(max(nB, 1), dtype=torch.uint8, device="cuda")
    x = sb.Isect.front_end(A, B, da.data_ptr(), db.data_ptr())
    ab, code = x.candidates()
    hab, seg = x.hits()
    r = x.job_ranges()
    fa, fb = da.cpu().numpy(), db.cpu().numpy()
    ta, tb = A.triangle_start.astype(np.int64), B.triangle_start.astype(np.int64)
    # every pair stays inside one job
    ja = np.searchsorted(ta, ab[:, 0], side="right") - 1
    jb = np.searchsorted(tb, ab[:, 1], side="right") - 1
    assert np.array_equal(ja, jb), "a candidate pair joins two different jobs"
    total_p = total_h = 0
    which = range(len(jobs)) if sample is None else sample
    for j in which:
        a, b = jobs[j]
        ref = oracle.candidate_pairs(a, b) if len(a[1]) and len(b[1]) else np.zeros((0, 2), np.uint32)
        mine = ab[ja == j].astype(np.int64) - [ta[j], tb[j]]
        assert np.array_equal(mine, ref.astype(np.int64)), "candidate pairs of job %d differ" % j
        if len(ref):
            ret, cop, hit, rseg = oracle.predicate_pairs(a, b, ref)
            assert np.array_equal(code[ja == j], (ret | (cop << 1)).astype(np.uint8)), "predicate codes of job %d" % j
            h = hit.astype(bool)
            hj = hab[r[j]:r[j + 1]].astype(np.int64) - [ta[j], tb[j]]
            assert np.array_equal(hj, ref[h].astype(np.int64)), "hit pairs of job %d" % j
            assert seg[r[j]:r[j + 1]].tobytes() == rseg[h].tobytes(), "segments of job %d" % j
            total_h += int(h.sum())
        else:
            assert r[j] == r[j + 1]
        total_p += len(ref)
        if len(a[1]):
            oa = oracle.classify(b, oracle.centroids(*a))[0] if len(b[1]) else np.zeros(len(a[1]), np.uint8)
            assert np.array_equal(fa[ta[j]:ta[j + 1]], oa), "inside flags of A, job %d" % j
        if len(b[1]):
            ob = oracle.classify(a, oracle.centroids(*b))[0] if len(a[1]) else np.zeros(len(b[1]), np.uint8)
            assert np.array_equal(fb[tb[j]:tb[j + 1]], ob), "inside flags of B, job %d" % j
    if sample is None:
        assert total_p == len(ab) and total_h == len(hab)
    # per-axis bits through the batch too (all three ray grids)
    pa = A.classify_faces_against(B)[1]
    for j in list(which)[:6]:
        a, b = jobs[j]
        if len(a[1]) and len(b[1]):
            assert np.array_equal(pa[ta[j]:ta[j + 1]], oracle.classify(b, oracle.centroids(*a))[1]), "per-axis bits, job %d" % j
    x.close(); A.close(); B.close()
    return total_p, total_h


def test_batch_ragged_jobs_vs_oracle(oracle):
    ctx = sb.Context(0)
    p, h = check_batch(ctx, oracle, ragged_jobs())
    assert p > 0 and h > 0
    ctx.close()


def test_batch_c5_shape_sampled(oracle):
    """125 jobs of the C5 shape (icosphere k=4 pairs: one rank's share of the 1,000-job batch);
    16 of them checked against the oracle."""
    ctx = sb.Context(0)
    jobs = [meshgen.config_c5_job(j) for j in range(125)]
    check_batch(ctx, oracle, jobs, sample=list(range(0, 125, 8)))
    ctx.close()


def test_batch_rebuild_and_single_job(oracle):
    """A batch of one job equals the plain mesh path; a rebuilt batch gives the same answer."""
    import torch
    ctx = sb.Context(0)
    a, b = meshgen.config_c5_job(7, k=3)
    pitch = sb.batch_pitch(a[0], b[0])
    A, B = sb.Mesh.batch(ctx, [a], pitch), sb.Mesh.batch(ctx, [b], pitch)
    ma, mb = ctx.mesh(*a), ctx.mesh(*b)
    x, y = A.intersect(B), ma.intersect(mb)
    assert np.array_equal(x.candidates()[0], y.candidates()[0]) and np.array_equal(x.hits()[0], y.hits()[0])
    assert x.hits()[1].tobytes() == y.hits()[1].tobytes()
    x.close()
    A.build(); B.build()
    x = A.intersect(B)
    assert np.array_equal(x.hits()[0], y.hits()[0])
    assert np.array_equal(A.classify_faces_against(B, per_axis=False)[0], ma.classify_faces_against(mb, per_axis=False)[0])
    x.close(); y.close(); A.close(); B.close(); ma.close(); mb.close()
    ctx.close()


def test_batch_bad_arguments():
    ctx = sb.Context(0)
    a, b = meshgen.icosphere(1), meshgen.icosphere(1, center=(0.3, 0, 0))
    with pytest.raises(sb.SolidBooleanError, match="lattice_pitch"):
        sb.Mesh.batch(ctx, [a], 0.0)
    with pytest.raises(sb.SolidBooleanError, match="too small"):
        sb.Mesh.batch(ctx, [a], 1.0).normals()
    bad = (a[0], np.array([[0, 1, 999]], np.uint32))
    with pytest.raises(sb.SolidBooleanError, match="out of range"):
        sb.Mesh.batch(ctx, [a, bad], 8.0).normals()
    A1 = sb.Mesh.batch(ctx, [a, a], 8.0)
    B1 = sb.Mesh.batch(ctx, [b], 8.0)
    B2 = sb.Mesh.batch(ctx, [b, b], 16.0)
    plain = ctx.mesh(*b)
    for other, msg in ((B1, "job counts"), (B2, "lattice pitches"), (plain, "another batch mesh")):
        with pytest.raises(sb.SolidBooleanError, match=msg):
            A1.intersect(other)
    with pytest.raises(sb.SolidBooleanError, match="belong to no job"):
        A1.classify(np.zeros((4, 3)))
    for m in (A1, B1, B2, plain):
        m.close()
    ctx.close()
