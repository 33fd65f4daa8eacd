"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle.

* ``Oracle``  -> oracle/libsboracle.so, the portable C restatement (sb_oracle.c).
* ``Ref``     -> oracle/_ref/libsbref.so, the unmodified reference compiled from
  /root/reference with oracle/ref_harness.cpp (present only where it was built).

Nothing under ``solidboolean_b200`` imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsbref.so")
HOOK_SO = os.path.join(HERE, "_ref", "libsbref_hook.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref)."""
    args = ["make", "-C", HERE, "all"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True, capture_output=True)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def fnv1a64(buf: bytes) -> int:
    h = 14695981039346656037
    # vectorising FNV is not possible; use the C helper when big
    lib = Oracle.get().lib
    arr = np.frombuffer(buf, dtype=np.uint8)
    return int(lib.sbo_fnv1a64(arr.ctypes.data_as(C.c_void_p), arr.size)) if arr.size else h


class Oracle:
    _inst = None

    @classmethod
    def get(cls) -> "Oracle":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.sbo_tri_tri_batch.argtypes = [_f64p, C.c_size_t, _i32p, _i32p, _f64p]
        L.sbo_normals.argtypes = [_f64p, _u32p, C.c_size_t, _f64p]
        L.sbo_tri_boxes.argtypes = [_f64p, _u32p, C.c_size_t, _f64p]
        L.sbo_centroids.argtypes = [_f64p, _u32p, C.c_size_t, _f64p]
        L.sbo_candidate_pairs.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint32))]
        L.sbo_candidate_pairs.restype = C.c_size_t
        L.sbo_free.argtypes = [C.c_void_p]
        L.sbo_predicate_pairs.argtypes = [_f64p, _u32p, _f64p, _u32p, _u32p, C.c_size_t,
                                          _i8p, _i8p, _u8p, _f64p]
        L.sbo_classify.argtypes = [_f64p, _u32p, C.c_size_t, _f64p, C.c_size_t, _u8p, _u8p,
                                   C.POINTER(C.c_uint64)]
        L.sbo_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.sbo_fnv1a64.restype = C.c_uint64
        L.sbo_uncut_half_edges.argtypes = [_u32p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_uint64, _u32p,
                                           C.POINTER(C.c_size_t), _u64p, _u32p, C.POINTER(C.c_size_t)]
        L.sbo_uncut_half_edges.restype = C.c_int
        L.sbo_uncut_adjacency.argtypes = [_u32p, _u32p, C.c_size_t, C.c_uint64, _u64p, _u32p, C.c_size_t, _i32p]
        L.sbo_uncut_components.argtypes = [_i32p, C.c_size_t, C.c_uint64, _u32p]
        L.sbo_uncut_components.restype = C.c_size_t
        L.sbo_face_groups.argtypes = [_u32p, C.c_size_t, _u64p, _u32p, C.c_size_t, _u32p, _u32p, C.c_size_t, C.c_size_t,
                                      C.c_size_t, _u32p, _u32p, C.POINTER(C.c_size_t)]
        L.sbo_face_groups.restype = C.c_size_t
        L.sbo_cut_contexts.argtypes = [_u32p, _f64p, C.c_size_t, C.c_int, _u32p, _u32p, _f64p, _u32p, _u32p]
        L.sbo_cut_contexts.restype = C.c_size_t

    # -- predicate ---------------------------------------------------------
    def tri_tri_batch(self, tris18):
        t = _f64(tris18).reshape(-1, 18)
        n = t.shape[0]
        ret = np.zeros(n, np.int32)
        cop = np.zeros(n, np.int32)
        seg = np.zeros((n, 6), np.float64)
        if n:
            self.lib.sbo_tri_tri_batch(t, n, ret, cop, seg)
        return ret, cop, seg

    # -- prepare() pieces --------------------------------------------------
    def normals(self, xyz, tri):
        xyz, tri = _f64(xyz), _u32(tri)
        out = np.zeros((tri.shape[0], 3), np.float64)
        self.lib.sbo_normals(xyz, tri, tri.shape[0], out)
        return out

    def tri_boxes(self, xyz, tri):
        xyz, tri = _f64(xyz), _u32(tri)
        out = np.zeros((tri.shape[0], 6), np.float64)
        self.lib.sbo_tri_boxes(xyz, tri, tri.shape[0], out)
        return out

    def centroids(self, xyz, tri):
        xyz, tri = _f64(xyz), _u32(tri)
        out = np.zeros((tri.shape[0], 3), np.float64)
        self.lib.sbo_centroids(xyz, tri, tri.shape[0], out)
        return out

    # -- broad + narrow phase ---------------------------------------------
    def candidate_pairs(self, meshA, meshB):
        """Sorted (a, b) pairs whose triangle boxes overlap. -> uint32 [P,2]"""
        ba = self.tri_boxes(*meshA)
        bb = self.tri_boxes(*meshB)
        ptr = C.POINTER(C.c_uint32)()
        n = self.lib.sbo_candidate_pairs(ba, ba.shape[0], bb, bb.shape[0], C.byref(ptr))
        if n:
            out = np.ctypeslib.as_array(ptr, shape=(n, 2)).copy()
        else:
            out = np.zeros((0, 2), np.uint32)
        self.lib.sbo_free(ptr)
        return out

    def predicate_pairs(self, meshA, meshB, pairs):
        xa, ta = _f64(meshA[0]), _u32(meshA[1])
        xb, tb = _f64(meshB[0]), _u32(meshB[1])
        pairs = _u32(pairs).reshape(-1, 2)
        n = pairs.shape[0]
        ret = np.zeros(n, np.int8)
        cop = np.zeros(n, np.int8)
        hit = np.zeros(n, np.uint8)
        seg = np.zeros((n, 6), np.float64)
        if n:
            self.lib.sbo_predicate_pairs(xa, ta, xb, tb, pairs, n, ret, cop, hit, seg)
        return ret, cop, hit, seg

    # -- uncut triangles + half-edge map -----------------------------------
    def uncut_half_edges(self, tri, cut=None, vertex_offset=0, triangle_offset=0):
        """addUnintersectedTriangles (src/solidboolean.cpp:250-286) for one mesh.
        -> dict(ok, face [n], keys [m] ascending, owner [m], adj [n,3])"""
        tri = _u32(tri).reshape(-1, 3)
        nT = tri.shape[0]
        cutp = None
        if cut is not None:
            cut = np.ascontiguousarray(cut, dtype=np.uint8)
            assert cut.shape[0] == nT
            cutp = cut.ctypes.data_as(C.c_void_p)
        face = np.zeros(max(nT, 1), np.uint32)
        keys = np.zeros(max(3 * nT, 1), np.uint64)
        owner = np.zeros(max(3 * nT, 1), np.uint32)
        n, m = C.c_size_t(0), C.c_size_t(0)
        ok = self.lib.sbo_uncut_half_edges(tri, nT, cutp, vertex_offset, triangle_offset, face, C.byref(n),
                                           keys, owner, C.byref(m))
        face, keys, owner = face[:n.value].copy(), keys[:m.value].copy(), owner[:m.value].copy()
        adj = np.full((n.value, 3), -1, np.int32)
        if n.value:
            self.lib.sbo_uncut_adjacency(tri, face, n.value, vertex_offset, keys, owner, m.value, adj)
        label = np.zeros(n.value, np.uint32)
        comps = self.lib.sbo_uncut_components(adj, n.value, triangle_offset, label) if n.value else 0
        return dict(ok=bool(ok), face=face, keys=keys, owner=owner, adj=adj, label=label, components=int(comps))

    # -- buildFaceGroups with loops -------------------------------------------
    def face_groups(self, tri, keys, owner, loops, remaining_start, remaining_count):
        """SolidBoolean::buildFaceGroups (src/solidboolean.cpp:167-239) on explicit inputs.
        loops: list of vertex cycles. -> group [nTri] (0xffffffff = unreached), order (assignment order), n_groups"""
        tri = _u32(tri).reshape(-1, 3)
        keys = np.ascontiguousarray(keys, np.uint64)
        owner = _u32(owner)
        ls = np.cumsum([0] + [len(l) for l in loops]).astype(np.uint32)
        lv = _u32(np.concatenate([np.asarray(l, np.uint32) for l in loops])) if loops else np.zeros(1, np.uint32)
        group = np.zeros(max(len(tri), 1), np.uint32)
        order = np.zeros(max(len(tri), 1), np.uint32)
        n = C.c_size_t(0)
        g = self.lib.sbo_face_groups(tri, len(tri), keys, owner, len(keys), ls, lv, len(loops), remaining_start,
                                     remaining_count, group, order, C.byref(n))
        return group[:len(tri)], order[:n.value], int(g)

    # -- per-triangle intersection contexts ---------------------------------
    def cut_contexts(self, hits, seg, which):
        """The pair-loop body of combine() (src/solidboolean.cpp:296-339) over `hits` in the given order.
        -> dict(tri [c], point_start [c+1], points [p,3], edge_start [c+1], edges [e,2])"""
        hits = _u32(hits).reshape(-1, 2)
        seg = _f64(seg).reshape(-1, 6)
        n = hits.shape[0]
        tri = np.zeros(max(n, 1), np.uint32)
        ps = np.zeros(n + 1, np.uint32)
        es = np.zeros(n + 1, np.uint32)
        pts = np.zeros((max(2 * n, 1), 3), np.float64)
        edges = np.zeros((max(n, 1), 2), np.uint32)
        c = self.lib.sbo_cut_contexts(hits, seg, n, which, tri, ps, pts, es, edges) if n else 0
        return dict(tri=tri[:c].copy(), point_start=ps[:c + 1].copy(), points=pts[:ps[c]].copy(),
                    edge_start=es[:c + 1].copy(), edges=edges[:es[c]].copy())

    # -- classification -----------------------------------------------------
    def classify(self, target_mesh, pts):
        xyz, tri = _f64(target_mesh[0]), _u32(target_mesh[1])
        pts = _f64(pts).reshape(-1, 3)
        q = pts.shape[0]
        inside = np.zeros(q, np.uint8)
        per_axis = np.zeros((q, 3), np.uint8)
        cand = C.c_uint64(0)
        if q:
            self.lib.sbo_classify(xyz, tri, tri.shape[0], pts, q, inside, per_axis, C.byref(cand))
        return inside, per_axis, int(cand.value)


class Ref:
    """The real reference (unmodified sources) behind oracle/ref_harness.cpp."""
    _inst = None

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    @classmethod
    def get(cls) -> "Ref":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_mesh_create.argtypes = [_f64p, C.c_size_t, _u32p, C.c_size_t]
        L.ref_mesh_create.restype = vp
        L.ref_mesh_prepare_ms.argtypes = [vp]
        L.ref_mesh_prepare_ms.restype = C.c_double
        L.ref_mesh_time_prepare.argtypes = [vp]
        L.ref_mesh_time_prepare.restype = C.c_double
        L.ref_mesh_normals.argtypes = [vp, _f64p]
        L.ref_mesh_boxes.argtypes = [vp, _f64p]
        L.ref_mesh_centroids.argtypes = [vp, _f64p]
        L.ref_mesh_destroy.argtypes = [vp]
        L.ref_op_create.argtypes = [vp, vp]
        L.ref_op_create.restype = vp
        L.ref_op_destroy.argtypes = [vp]
        L.ref_op_search.argtypes = [vp, C.POINTER(C.c_double)]
        L.ref_op_search.restype = C.c_size_t
        L.ref_op_pairs.argtypes = [vp, _u32p]
        L.ref_op_predicate.argtypes = [vp, _u32p, C.c_size_t, _i8p, _i8p, _u8p, _f64p]
        L.ref_op_predicate.restype = C.c_double
        L.ref_tri_tri_batch.argtypes = [_f64p, C.c_size_t, _i32p, _i32p, _f64p]
        L.ref_op_classify.argtypes = [vp, C.c_int, _f64p, C.c_size_t, _u8p, _u8p]
        L.ref_op_classify.restype = C.c_double
        L.ref_op_combine.argtypes = [vp, _f64p, C.c_int]
        L.ref_op_combine.restype = C.c_int
        L.ref_op_result_vertex_count.argtypes = [vp]
        L.ref_op_result_vertex_count.restype = C.c_size_t
        L.ref_op_result_vertices.argtypes = [vp, _f64p]
        L.ref_op_result_triangle_count.argtypes = [vp, C.c_int]
        L.ref_op_result_triangle_count.restype = C.c_size_t
        L.ref_op_result_triangles.argtypes = [vp, C.c_int, _u32p]
        L.ref_op_group_count.argtypes = [vp, C.c_int]
        L.ref_op_group_count.restype = C.c_size_t
        if hasattr(L, "ref_op_uncut"):
            L.ref_op_uncut.argtypes = [vp, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int]
            L.ref_op_uncut.restype = C.c_int
            L.ref_op_uncut_fetch.argtypes = [vp, C.c_int, _u64p, _u32p]
            L.ref_op_uncut_lookup.argtypes = [vp, C.c_int, _u64p, C.c_size_t, _i32p]
            L.ref_op_uncut_groups.argtypes = [vp, C.c_int, _u32p]
            L.ref_op_uncut_groups.restype = C.c_size_t
            L.ref_face_groups.argtypes = [_u32p, C.c_size_t, _u64p, _u32p, C.c_size_t, _u32p, _u32p, C.c_size_t, C.c_size_t,
                                          C.c_size_t, _u32p, _u32p]
            L.ref_face_groups.restype = C.c_size_t
            L.ref_op_uncut_ms.argtypes = [vp, C.c_int, C.c_int]
            L.ref_op_uncut_ms.restype = C.c_double
        if hasattr(L, "ref_load_obj"):
            L.ref_load_obj.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p,
                                       C.POINTER(C.c_size_t)]
            L.ref_load_obj.restype = C.c_int

    def tri_tri_batch(self, tris18):
        t = _f64(tris18).reshape(-1, 18)
        n = t.shape[0]
        ret = np.zeros(n, np.int32)
        cop = np.zeros(n, np.int32)
        seg = np.zeros((n, 6), np.float64)
        if n:
            self.lib.ref_tri_tri_batch(t, n, ret, cop, seg)
        return ret, cop, seg

    def load_obj(self, path):
        nv, nt = C.c_size_t(0), C.c_size_t(0)
        if not self.lib.ref_load_obj(path.encode(), None, C.byref(nv), None, C.byref(nt)):
            raise IOError(path)
        xyz = np.zeros((nv.value, 3), np.float64)
        tri = np.zeros((nt.value, 3), np.uint32)
        self.lib.ref_load_obj(path.encode(), xyz.ctypes.data_as(C.c_void_p), C.byref(nv),
                              tri.ctypes.data_as(C.c_void_p), C.byref(nt))
        return xyz, tri

    def face_groups(self, tri, keys, owner, loops, remaining_start, remaining_count):
        """The reference's private buildFaceGroups on explicit inputs (same signature as Oracle.face_groups)."""
        tri = _u32(tri).reshape(-1, 3)
        keys = np.ascontiguousarray(keys, np.uint64)
        owner = _u32(owner)
        ls = np.cumsum([0] + [len(l) for l in loops]).astype(np.uint32)
        lv = _u32(np.concatenate([np.asarray(l, np.uint32) for l in loops])) if loops else np.zeros(1, np.uint32)
        group = np.zeros(max(len(tri), 1), np.uint32)
        order = np.zeros(max(len(tri), 1), np.uint32)
        g = self.lib.ref_face_groups(tri, len(tri), keys, owner, len(keys), ls, lv, len(loops), remaining_start,
                                     remaining_count, group, order)
        n = int((group[:len(tri)] != 0xffffffff).sum())
        return group[:len(tri)], order[:n], int(g)

    def mesh(self, xyz, tri) -> "RefMesh":
        return RefMesh(self, xyz, tri)

    def op(self, a: "RefMesh", b: "RefMesh") -> "RefOp":
        return RefOp(self, a, b)


class RefMesh:
    def __init__(self, ref: Ref, xyz, tri):
        self.ref = ref
        self.xyz = _f64(xyz).reshape(-1, 3)
        self.tri = _u32(tri).reshape(-1, 3)
        self.h = ref.lib.ref_mesh_create(self.xyz, self.xyz.shape[0], self.tri, self.tri.shape[0])

    @property
    def prepare_ms(self):
        return self.ref.lib.ref_mesh_prepare_ms(self.h)

    def time_prepare(self):
        return self.ref.lib.ref_mesh_time_prepare(self.h)

    def normals(self):
        out = np.zeros((self.tri.shape[0], 3), np.float64)
        self.ref.lib.ref_mesh_normals(self.h, out)
        return out

    def boxes(self):
        out = np.zeros((self.tri.shape[0], 6), np.float64)
        self.ref.lib.ref_mesh_boxes(self.h, out)
        return out

    def centroids(self):
        out = np.zeros((self.tri.shape[0], 3), np.float64)
        self.ref.lib.ref_mesh_centroids(self.h, out)
        return out

    def close(self):
        if self.h:
            self.ref.lib.ref_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefOp:
    def __init__(self, ref: Ref, a: RefMesh, b: RefMesh):
        self.ref, self.a, self.b = ref, a, b
        self.h = ref.lib.ref_op_create(a.h, b.h)
        self.search_ms = None

    def search(self):
        """Reference broad phase. -> pairs uint32 [P,2] in the reference's order."""
        ms = C.c_double(0)
        n = self.ref.lib.ref_op_search(self.h, C.byref(ms))
        self.search_ms = ms.value
        out = np.zeros((n, 2), np.uint32)
        if n:
            self.ref.lib.ref_op_pairs(self.h, out)
        return out

    def predicate(self, pairs):
        pairs = _u32(pairs).reshape(-1, 2)
        n = pairs.shape[0]
        ret = np.zeros(n, np.int8)
        cop = np.zeros(n, np.int8)
        hit = np.zeros(n, np.uint8)
        seg = np.zeros((n, 6), np.float64)
        ms = self.ref.lib.ref_op_predicate(self.h, pairs, n, ret, cop, hit, seg) if n else 0.0
        return ret, cop, hit, seg, ms

    def classify(self, target: int, pts):
        pts = _f64(pts).reshape(-1, 3)
        q = pts.shape[0]
        inside = np.zeros(q, np.uint8)
        per_axis = np.zeros((q, 3), np.uint8)
        ms = self.ref.lib.ref_op_classify(self.h, target, pts, q, inside, per_axis) if q else 0.0
        return inside, per_axis, ms

    def combine(self, quiet=True):
        stage = np.zeros(7, np.float64)
        ok = bool(self.ref.lib.ref_op_combine(self.h, stage, 1 if quiet else 0))
        res = {"ok": ok, "stage_ms": stage}
        if ok:
            nv = self.ref.lib.ref_op_result_vertex_count(self.h)
            v = np.zeros((nv, 3), np.float64)
            self.ref.lib.ref_op_result_vertices(self.h, v)
            res["vertices"] = v
            for which, name in enumerate(("union", "diff", "intersect")):
                nt = self.ref.lib.ref_op_result_triangle_count(self.h, which)
                t = np.zeros((nt, 3), np.uint32)
                if nt:
                    self.ref.lib.ref_op_result_triangles(self.h, which, t)
                res[name] = t
            res["groups"] = (self.ref.lib.ref_op_group_count(self.h, 0), self.ref.lib.ref_op_group_count(self.h, 1))
        return res

    def uncut(self, cut_a=None, cut_b=None):
        """The reference's addUnintersectedTriangles for both meshes, as combine() calls them.
        -> [dict(ok, n_triangles (m_newTriangles.size() after the call), keys, owner)] x 2, triangles"""
        def ptr(c):
            if c is None:
                return None, None
            c = np.ascontiguousarray(c, dtype=np.uint8)
            return c, c.ctypes.data_as(C.c_void_p)
        ka, pa = ptr(cut_a)
        kb, pb = ptr(cut_b)
        ntri = (C.c_size_t * 2)()
        nkeys = (C.c_size_t * 2)()
        res = self.ref.lib.ref_op_uncut(self.h, pa, pb, ntri, nkeys, 1)
        out = []
        for w in range(2):
            keys = np.zeros(max(nkeys[w], 1), np.uint64)
            owner = np.zeros(max(nkeys[w], 1), np.uint32)
            self.ref.lib.ref_op_uncut_fetch(self.h, w, keys, owner)
            out.append(dict(ok=bool(res >> w & 1), n_triangles=int(ntri[w]), keys=keys[:nkeys[w]], owner=owner[:nkeys[w]]))
        tris = np.zeros((max(ntri[1], 1), 3), np.uint32)
        if ntri[1]:
            self.ref.lib.ref_op_result_triangles(self.h, 0, tris)
        return out, tris[:ntri[1]]

    def uncut_groups(self, which, n):
        """buildFaceGroups without loops over the state left by uncut(): -> label [n], group count"""
        label = np.zeros(max(n, 1), np.uint32)
        g = self.ref.lib.ref_op_uncut_groups(self.h, which, label)
        return label[:n], int(g)

    def uncut_ms(self, what, which):
        """ms inside the reference's addUnintersectedTriangles (what=0) / buildFaceGroups (what=1)"""
        return float(self.ref.lib.ref_op_uncut_ms(self.h, what, which))

    def uncut_lookup(self, which, from_to):
        ft = np.ascontiguousarray(from_to, dtype=np.uint64).reshape(-1, 2)
        out = np.zeros(ft.shape[0], np.int32)
        if ft.shape[0]:
            self.ref.lib.ref_op_uncut_lookup(self.h, which, ft, ft.shape[0], out)
        return out

    def close(self):
        if self.h:
            self.ref.lib.ref_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
