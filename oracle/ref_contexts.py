"""TEST INFRASTRUCTURE ONLY: capture the IntersectedContext objects the UNMODIFIED reference builds
inside SolidBoolean::combine() (src/solidboolean.cpp:296-339) with the observation hook
oracle/ref_hook.cpp.  Must run in a process of its own with the hook preloaded:

    LD_PRELOAD=oracle/_ref/libsbref_hook.so SBREF_PATH=oracle/_ref/libsbref.so \\
        python -m oracle.ref_contexts in.npz out.npz [--sorted]

in.npz: xyz_a, tri_a, xyz_b, tri_b.  out.npz: per side s in (a, b): s_tri (triangle id per context,
ascending), s_point_start, s_points, s_edge_start, s_edges -- the layout of Oracle.cut_contexts --
plus ok (combine()'s return value) and the hit list hits / seg in the order the loop saw it.
`capture()` is the wrapper the tests and the fixture generator call."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def capture(a, b, sorted_pairs=True):
    """Run this module in a subprocess with the hook preloaded. -> dict of the out.npz arrays"""
    from oracle import HOOK_SO, REF_SO
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.npz"), os.path.join(tmp, "out.npz")
        np.savez(fin, xyz_a=a[0], tri_a=a[1], xyz_b=b[0], tri_b=b[1])
        env = dict(os.environ, LD_PRELOAD=HOOK_SO, SBREF_PATH=REF_SO, PYTHONPATH=ROOT)
        cmd = [sys.executable, "-m", "oracle.ref_contexts", fin, fout] + (["--sorted"] if sorted_pairs else [])
        r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("ref_contexts failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
        with np.load(fout) as z:
            return {k: z[k] for k in z.files}


def main():
    fin, fout = sys.argv[1], sys.argv[2]
    want_sorted = "--sorted" in sys.argv
    hook = C.CDLL(os.environ["LD_PRELOAD"])      # already mapped: same handle
    hook.hook_count.restype = C.c_size_t
    hook.hook_sizes.argtypes = [C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    hook.hook_get.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    from oracle import Ref
    R = Ref.get()
    z = np.load(fin)
    a = (np.ascontiguousarray(z["xyz_a"], np.float64), np.ascontiguousarray(z["tri_a"], np.uint32))
    b = (np.ascontiguousarray(z["xyz_b"], np.float64), np.ascontiguousarray(z["tri_b"], np.uint32))
    ma, mb = R.mesh(*a), R.mesh(*b)
    op = R.op(ma, mb)
    # the pair list exactly as the loop will see it (the wrapped search sorts it on request)
    hook.hook_reset(1 if want_sorted else 0)
    pairs = op.search()
    ret, cop, hit, seg, _ = op.predicate(pairs)
    hits, hseg = pairs[hit.astype(bool)], seg[hit.astype(bool)]
    hook.hook_reset(1 if want_sorted else 0)
    res = op.combine()
    n = hook.hook_count()
    out = {"ok": np.array([1 if res["ok"] else 0], np.uint8), "hits": hits, "seg": hseg}
    # which triangle a captured context belongs to: the retriangulator was constructed from the
    # triangle's vertices and normal (src/solidboolean.cpp:358-365) -> projection origin = vertex 0,
    # projection axis = unit(vertex 1 - vertex 0) (src/retriangulator.cpp:27-37)
    sides = ((a, ma.normals(), np.unique(hits[:, 0])), (b, mb.normals(), np.unique(hits[:, 1])))
    per_side = ([], [])
    n_first = len(sides[0][2])
    for i in range(n):
        npts, ned = C.c_size_t(0), C.c_size_t(0)
        hook.hook_sizes(i, C.byref(npts), C.byref(ned))
        origin, normal, axis = np.zeros(3), np.zeros(3), np.zeros(3)
        pts = np.zeros((max(npts.value, 1), 3))
        edges = np.zeros((max(ned.value, 1), 2), np.uint32)
        hook.hook_get(i, origin.ctypes.data, normal.ctypes.data, axis.ctypes.data, pts.ctypes.data, edges.ctypes.data)
        s = 0 if i < n_first else 1
        (xyz, tri), normals, cut = sides[s]
        match = [t for t in cut if np.array_equal(xyz[tri[t, 0]], origin) and np.array_equal(normals[t], normal)]
        if len(match) > 1:   # coplanar triangles sharing their first vertex: the projection axis = unit(v1 - v0) tells them apart
            def unit(v):
                return v / np.linalg.norm(v)
            match = [t for t in match if np.allclose(unit(xyz[tri[t, 1]] - xyz[tri[t, 0]]), axis, rtol=0, atol=1e-9)]
        assert len(match) == 1, (i, s, match)
        per_side[s].append((int(match[0]), pts[:npts.value], edges[:ned.value]))
    for s, name in enumerate("ab"):
        ctx = sorted(per_side[s], key=lambda c: c[0])
        out[name + "_tri"] = np.array([c[0] for c in ctx], np.uint32)
        out[name + "_point_start"] = np.cumsum([0] + [len(c[1]) for c in ctx]).astype(np.uint32)
        out[name + "_points"] = np.concatenate([c[1] for c in ctx]) if ctx else np.zeros((0, 3))
        out[name + "_edge_start"] = np.cumsum([0] + [len(c[2]) for c in ctx]).astype(np.uint32)
        out[name + "_edges"] = np.concatenate([c[2] for c in ctx]) if ctx else np.zeros((0, 2), np.uint32)
    np.savez(fout, **out)


if __name__ == "__main__":
    main()
