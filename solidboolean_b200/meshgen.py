"""Deterministic synthetic closed meshes for the benchmark configs.

All generators return ``(xyz float64 [nV,3], tri uint32 [nT,3])`` with outward
(counter-clockwise seen from outside) winding.  No RNG is used for C2-C4; C5
seeds its offsets from the job id.  ``round_to_float=True`` rounds coordinates
to float32 first (what an OBJ loaded through tinyobj looks like,
reference test/main.cpp:51-57).
"""
from __future__ import annotations

import numpy as np


def _finish(xyz, tri, round_to_float):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if round_to_float:
        xyz = xyz.astype(np.float32).astype(np.float64)
    return xyz, np.ascontiguousarray(tri, dtype=np.uint32)


def icosphere(k: int, radius: float = 1.0, center=(0.0, 0.0, 0.0), round_to_float: bool = False):
    """Golden-ratio icosahedron, k midpoint subdivisions: 20*4^k triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([
        [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
        [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
        [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(k):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a = uniq // nv
        b = uniq % nv
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = f.shape[0]
        m01 = nv + inv[:nf]
        m12 = nv + inv[nf:2 * nf]
        m20 = nv + inv[2 * nf:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], axis=1),
            np.stack([f[:, 1], m12, m01], axis=1),
            np.stack([f[:, 2], m20, m12], axis=1),
            np.stack([m01, m12, m20], axis=1)], axis=0)
    xyz = v * radius + np.asarray(center, dtype=np.float64)
    return _finish(xyz, f, round_to_float)


def torus(nu: int, nv: int, R: float = 1.0, r: float = 0.35, center=(0.0, 0.0, 0.0),
          round_to_float: bool = False):
    """Torus around the z axis, nu x nv quads split into 2*nu*nv triangles."""
    u = np.arange(nu, dtype=np.float64) * (2.0 * np.pi / nu)
    w = np.arange(nv, dtype=np.float64) * (2.0 * np.pi / nv)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    x = (R + r * np.cos(ww)) * np.cos(uu)
    y = (R + r * np.cos(ww)) * np.sin(uu)
    z = r * np.sin(ww)
    xyz = np.stack([x, y, z], axis=-1).reshape(-1, 3) + np.asarray(center, dtype=np.float64)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    i1 = (i + 1) % nu
    j1 = (j + 1) % nv
    a = (i * nv + j).ravel()
    b = (i1 * nv + j).ravel()
    c = (i1 * nv + j1).ravel()
    d = (i * nv + j1).ravel()
    tri = np.concatenate([np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)], axis=0)
    return _finish(xyz, tri, round_to_float)


def slab(n: int, size: float = 1.0, thickness: float = 0.05, center=(0.0, 0.0, 0.0),
         tilt: float = 0.0, round_to_float: bool = False):
    """Closed tessellated slab: n x n quads on the top and bottom faces, n x 1
    quads on each of the four rims -> 4*n*n + 8*n triangles.  `tilt` rotates the
    slab about the x axis (radians).  Used for the dense-pair config C4."""
    g = np.linspace(-0.5 * size, 0.5 * size, n + 1)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    top = np.stack([gx, gy, np.full_like(gx, 0.5 * thickness)], axis=-1).reshape(-1, 3)
    bot = np.stack([gx, gy, np.full_like(gx, -0.5 * thickness)], axis=-1).reshape(-1, 3)
    xyz = np.concatenate([top, bot], axis=0)
    m = n + 1
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    a = (i * m + j).ravel()
    b = ((i + 1) * m + j).ravel()
    c = ((i + 1) * m + j + 1).ravel()
    d = (i * m + j + 1).ravel()
    tris = [np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)]
    off = m * m
    tris += [np.stack([a + off, c + off, b + off], axis=1), np.stack([a + off, d + off, c + off], axis=1)]
    k = np.arange(n)

    def rim(t0, t1):
        # t0 -> t1 is a top-edge direction such that (t0, b0, b1) faces outward
        b0, b1 = t0 + off, t1 + off
        return [np.stack([t0, b0, b1], axis=1), np.stack([t0, b1, t1], axis=1)]

    tris += rim(k * m, (k + 1) * m)                          # y = -size/2 side
    tris += rim((k + 1) * m + n, k * m + n)                  # y = +size/2 side
    tris += rim(k + 1, k)                                    # x = -size/2 side
    tris += rim(n * m + k, n * m + k + 1)                    # x = +size/2 side
    tri = np.concatenate(tris, axis=0)
    if tilt != 0.0:
        cs, sn = np.cos(tilt), np.sin(tilt)
        rot = np.array([[1, 0, 0], [0, cs, -sn], [0, sn, cs]], dtype=np.float64)
        xyz = xyz @ rot.T
    xyz = xyz + np.asarray(center, dtype=np.float64)
    return _finish(xyz, tri, round_to_float)


def signed_volume(xyz, tri) -> float:
    a = xyz[tri[:, 0].astype(np.int64)]
    b = xyz[tri[:, 1].astype(np.int64)]
    c = xyz[tri[:, 2].astype(np.int64)]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def is_closed_manifold(tri) -> bool:
    """Every directed edge appears once and its reverse appears once."""
    t = tri.astype(np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
    n = int(t.max()) + 1
    fwd = np.sort(e[:, 0] * n + e[:, 1])
    rev = np.sort(e[:, 1] * n + e[:, 0])
    return bool(np.all(np.diff(fwd) > 0) and np.array_equal(fwd, rev))


# ---- named benchmark configs (SURVEY 8 "Config sizes") ----------------------

def config_c2(round_to_float: bool = False):
    """Two offset icospheres, k=6: 81,920 + 81,920 triangles."""
    a = icosphere(6, round_to_float=round_to_float)
    b = icosphere(6, center=(0.71, 0.13, 0.07), round_to_float=round_to_float)
    return a, b


def config_c3(round_to_float: bool = False):
    """Icosphere k=8 (1,310,720 tris) vs torus 1024x512 (1,048,576 tris)."""
    a = icosphere(8, round_to_float=round_to_float)
    b = torus(1024, 512, R=1.0, r=0.35, center=(0.013, 0.007, 0.011), round_to_float=round_to_float)
    return a, b


def config_c4(k: int = 7, offset: float = 1.5e-3, round_to_float: bool = False):
    """Dense candidate-pair proxy: near-coincident icospheres."""
    a = icosphere(k, round_to_float=round_to_float)
    b = icosphere(k, center=(offset, 0.4 * offset, 0.2 * offset), round_to_float=round_to_float)
    return a, b


def config_c5_job(job_id: int, k: int = 4, round_to_float: bool = True):
    """One job of the batch config: icosphere pair with a seeded offset."""
    rng = np.random.default_rng(job_id)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    off = d * rng.uniform(0.3, 1.2)
    a = icosphere(k, round_to_float=round_to_float)
    b = icosphere(k, center=tuple(off), round_to_float=round_to_float)
    return a, b
