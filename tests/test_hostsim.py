"""CPU: the device predicate / ray headers compiled for the host (-DSB_HOST_SIM)
agree with the reference fixtures.  This checks the LOGIC of the device code
(permutation tables, branch structure, the sign filter) before GPU time is
spent; the -m gpu tests check the real thing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
SO = os.path.join(HERE, "hostsim", "libhostsim.so")


@pytest.fixture(scope="module")
def hs():
    deps = [SRC] + [os.path.join(HERE, "..", "solidboolean_b200", "csrc", f)
                    for f in ("sb_fp64.cuh", "sb_tritri.cuh", "sb_raytri.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-DSB_HOST_SIM", "-shared", "-o", SO, SRC],
                       check=True, capture_output=True)
    return C.CDLL(SO)


def run_tritri(hs, tris):
    tris = np.ascontiguousarray(tris, np.float64)
    n = len(tris)
    ret = np.zeros(n, np.int32); cop = np.zeros(n, np.int32); seg = np.zeros((n, 6))
    vp = C.c_void_p
    hs.hs_tri_tri_batch(tris.ctypes.data_as(vp), C.c_size_t(n), ret.ctypes.data_as(vp), cop.ctypes.data_as(vp),
                        seg.ctypes.data_as(vp))
    return ret, cop, seg


def test_device_predicate_logic_matches_reference_fixture(hs, golden_kat):
    ret, cop, seg = run_tritri(hs, golden_kat["tris"])
    assert np.array_equal(ret, golden_kat["ret"]) and np.array_equal(cop, golden_kat["coplanar"])
    assert seg.tobytes() == golden_kat["seg"].tobytes()


def test_device_predicate_logic_matches_oracle_random(hs, oracle):
    rng = np.random.default_rng(21)
    tris = np.concatenate([rng.uniform(-1, 1, (100000, 18)), rng.integers(-2, 3, (100000, 18)).astype(np.float64)])
    r0, c0, s0 = oracle.tri_tri_batch(tris)
    r1, c1, s1 = run_tritri(hs, tris)
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and s0.tobytes() == s1.tobytes()


def ray_cases(rng, n):
    """(point, triangle) records biased towards the degenerate cases that decide
    parity: rays through edges and vertices, grazing planes, tiny triangles."""
    t = rng.uniform(-1, 1, (n, 3, 3))
    kind = rng.integers(0, 8, n)
    axis = rng.integers(0, 3, n).astype(np.int32)
    w = rng.dirichlet([1, 1, 1], n)
    inside_pt = np.einsum("nk,nkd->nd", w, t)                     # a point of the triangle
    edge_pt = t[:, 0] + (t[:, 1] - t[:, 0]) * rng.uniform(0, 1, (n, 1))
    vert_pt = t[np.arange(n), rng.integers(0, 3, n)]
    target = np.where((kind == 0)[:, None], edge_pt, np.where((kind == 1)[:, None], vert_pt, inside_pt))
    target = np.where((kind == 2)[:, None], rng.uniform(-1, 1, (n, 3)), target)
    p = target.copy()
    back = rng.uniform(0.0, 2.0, n)
    p[np.arange(n), axis] -= back                                   # start behind the target along the ray axis
    p[kind == 3] = target[kind == 3]                                # start ON the triangle
    tiny = kind == 4
    t[tiny] = t[tiny, :1] + (t[tiny] - t[tiny, :1]) * 1e-9          # tiny triangles
    flat = kind == 5
    t[flat, :, 0] = t[flat, :1, 0]                                  # triangle parallel to x rays
    f32 = kind == 6
    t[f32] = t[f32].astype(np.float32); p[f32] = p[f32].astype(np.float32)
    grid = kind == 7
    t[grid] = np.round(t[grid] * 4) / 4; p[grid] = np.round(p[grid] * 4) / 4
    return np.ascontiguousarray(np.concatenate([p[:, None, :], t], axis=1).reshape(n, 12)), axis


def run_ray(hs, rec, axis, filtered):
    n = len(rec)
    flag = np.zeros(n, np.uint8)
    keys = np.zeros((n, 3), np.int64)
    vp = C.c_void_p
    hs.hs_ray_tri_batch(rec.ctypes.data_as(vp), axis.ctypes.data_as(vp), C.c_size_t(n), flag.ctypes.data_as(vp),
                        keys.ctypes.data_as(vp), C.c_int(filtered))
    return flag, keys


def test_sign_filter_equals_exact_sequence(hs):
    rng = np.random.default_rng(33)
    rec, axis = ray_cases(rng, 1500000)
    f0, k0 = run_ray(hs, rec, axis, 0)
    f1, k1 = run_ray(hs, rec, axis, 1)
    assert 0.1 < f0.mean() < 0.9
    assert np.array_equal(f0, f1) and np.array_equal(k0, k1)


def test_ray_triangle_matches_oracle(hs, oracle):
    """Exact device sequence vs the oracle's isPointInMesh on one-triangle meshes."""
    rng = np.random.default_rng(4)
    rec, axis = ray_cases(rng, 3000)
    flag, _ = run_ray(hs, rec, axis, 0)
    tri = np.array([[0, 1, 2]], np.uint32)
    for i in range(len(rec)):
        xyz = rec[i, 3:].reshape(3, 3)
        _, per, _ = oracle.classify((xyz, tri), rec[i, :3][None])
        assert per[0, axis[i]] == flag[i], i


def test_subnormal_ray_division_equals_ieee_division(tmp_path):
    """xdiv_huge_den (sb_raytri.cuh): scaled division + exact tie repair == IEEE division, bit for bit, on 5 M random operand
    pairs and 3 M constructed ties (tests/hostsim/div_huge_den.c; 1.6e8 pairs were run once when it was written)."""
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim", "div_huge_den.c")
    exe = str(tmp_path / "div_huge_den")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True, capture_output=True)
    r = subprocess.run([exe, "5000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:]
    assert "mismatches 0" in r.stdout
    fixes = int(r.stdout.strip().split("fixes")[-1])
    assert fixes > 1000          # the tie repair was exercised
