"""solidboolean_b200 -- B200 (sm_100a) intersection front end of solidboolean.

This package is a thin ctypes binding over the C ABI declared in
``include/solidboolean_b200.h`` (the drop-in boundary; the C++ SolidMesh /
SolidBoolean classes in ``solidboolean_b200/host`` sit on the same ABI).  There
is no CPU implementation here: importing works anywhere, but every compute call
needs the CUDA library ``lib/libsolidboolean_b200.so`` and a GPU, and raises
``SolidBooleanError`` otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB_PATH") or os.path.join(HERE, "lib", "libsolidboolean_b200.so")  # override: dev builds

STAGES = ("build", "broad", "narrow", "classify", "predicate", "halfedge", "contexts", "shard")
ISECT_NO_SORT = 1


class SolidBooleanError(RuntimeError):
    pass


class _BvhInfo(C.Structure):
    _fields_ = [("cluster_size", C.c_uint32), ("num_clusters", C.c_uint32),
                ("num_internal", C.c_uint32), ("root", C.c_int32)]


class _CommRankInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("selected_a", C.c_size_t), ("selected_b", C.c_size_t), ("z_lo", C.c_double),
                ("z_hi", C.c_double), ("candidates", C.c_size_t), ("hits", C.c_size_t), ("fallbacks", C.c_uint64)]


class _GridInfo(C.Structure):
    _fields_ = [("nu", C.c_uint32 * 3), ("nv", C.c_uint32 * 3), ("total_cells", C.c_uint32),
                ("total_refs", C.c_uint32), ("big", C.c_uint32 * 3), ("mean_extent", C.c_float * 3)]


_lib = None

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks
# against the header.
_vp = C.c_void_p
_sz = C.c_size_t
ABI = {
    "sb_last_error": (C.c_char_p, []),
    "sb_version": (C.c_char_p, []),
    "sb_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "sb_context_destroy": (None, [_vp]),
    "sb_context_synchronize": (C.c_int, [_vp]),
    "sb_context_stream": (_vp, [_vp]),
    "sb_context_device": (C.c_int, [_vp]),
    "sb_mesh_create": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_vp)]),
    "sb_mesh_upload": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_vp)]),
    "sb_batch_upload": (C.c_int, [_vp, _sz, _vp, _vp, _vp, _vp, C.c_double, C.POINTER(_vp)]),
    "sb_batch_info": (C.c_int, [_vp, C.POINTER(_sz), _vp, _vp]),
    "sb_batch_job_ranges": (C.c_int, [_vp, _vp]),
    "sb_shard_create": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "sb_shard_front_end": (C.c_int, [_vp, C.c_uint, C.POINTER(_vp), _vp, _vp]),
    "sb_shard_info": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_uint64)]),
    "sb_shard_destroy": (None, [_vp]),
    "sb_comm_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]),
    "sb_comm_size": (C.c_int, [_vp]),
    "sb_comm_set_meshes": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "sb_comm_front_end": (C.c_int, [_vp, C.c_uint, C.POINTER(_sz), C.POINTER(_sz), _vp, _vp]),
    "sb_comm_hits": (C.c_int, [_vp, _vp, _vp]),
    "sb_comm_rank": (C.c_int, [_vp, C.c_int, C.POINTER(_CommRankInfo)]),
    "sb_comm_destroy": (None, [_vp]),
    "sb_mesh_upload_device": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_vp)]),
    "sb_mesh_update": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "sb_mesh_build": (C.c_int, [_vp]),
    "sb_mesh_destroy": (None, [_vp]),
    "sb_mesh_num_triangles": (_sz, [_vp]),
    "sb_mesh_num_vertices": (_sz, [_vp]),
    "sb_mesh_normals": (C.c_int, [_vp, _vp]),
    "sb_mesh_triangle_boxes": (C.c_int, [_vp, _vp]),
    "sb_mesh_bounds": (C.c_int, [_vp, _vp]),
    "sb_mesh_order": (C.c_int, [_vp, _vp]),
    "sb_mesh_bvh_info": (C.c_int, [_vp, C.POINTER(_BvhInfo)]),
    "sb_mesh_bvh_nodes": (C.c_int, [_vp, _vp]),
    "sb_mesh_bvh_leaves": (C.c_int, [_vp, _vp, C.POINTER(_sz)]),
    "sb_mesh_grid_info": (C.c_int, [_vp, C.POINTER(_GridInfo)]),
    "sb_intersect": (C.c_int, [_vp, _vp, C.c_uint, C.POINTER(_vp)]),
    "sb_intersect_range": (C.c_int, [_vp, _vp, _sz, _sz, C.c_uint, C.POINTER(_vp)]),
    "sb_isect_destroy": (None, [_vp]),
    "sb_isect_counts": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "sb_isect_candidates": (C.c_int, [_vp, _vp, _vp]),
    "sb_isect_hits": (C.c_int, [_vp, _vp, _vp]),
    "sb_isect_hit_edges": (C.c_int, [_vp, _vp]),
    "sb_isect_face_flags": (C.c_int, [_vp, _vp, _vp]),
    "sb_isect_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_uint), C.POINTER(_vp),
                                       C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "sb_isect_pack_device": (C.c_int, [_vp, _vp, _sz]),
    "sb_isect_path_counts": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "sb_fp64_peak": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sb_tri_tri_batch": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "sb_isect_contexts": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "sb_cuts_destroy": (None, [_vp]),
    "sb_cuts_counts": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "sb_cuts_fetch": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_cuts_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "sb_isect_uncut": (C.c_int, [_vp, C.c_int, _sz, _sz, C.POINTER(_vp)]),
    "sb_mesh_uncut": (C.c_int, [_vp, _vp, _sz, _sz, C.POINTER(_vp)]),
    "sb_uncut_destroy": (None, [_vp]),
    "sb_uncut_counts": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(C.c_int)]),
    "sb_uncut_triangles": (C.c_int, [_vp, _vp, _vp]),
    "sb_uncut_half_edges": (C.c_int, [_vp, _vp, _vp]),
    "sb_uncut_adjacency": (C.c_int, [_vp, _vp]),
    "sb_uncut_components": (C.c_int, [_vp, _vp, C.POINTER(_sz)]),
    "sb_uncut_face_groups": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, C.POINTER(_sz)]),
    "sb_uncut_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                       C.POINTER(_vp)]),
    "sb_classify": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "sb_classify_faces": (C.c_int, [_vp, _vp, _vp, _vp]),
    "sb_classify_faces_device": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "sb_front_end": (C.c_int, [_vp, _vp, C.c_uint, C.POINTER(_vp), _vp, _vp]),
    "sb_front_end_range": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, C.c_uint, C.POINTER(_vp), _vp, _vp]),
    "sb_front_end_host": (C.c_int, [_vp, _vp, C.c_uint, C.POINTER(_vp), _vp, _vp, _vp, _vp, _sz]),
    "sb_context_enable_timing": (C.c_int, [_vp, C.c_int]),
    "sb_context_reset_timing": (C.c_int, [_vp]),
    "sb_context_get_timing": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "sb_context_classify_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}


def load_library():
    """Load the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SolidBooleanError(
            "CUDA library %s is missing: run `python -m solidboolean_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(code):
    if code != 0:
        msg = load_library().sb_last_error().decode(errors="replace")
        raise SolidBooleanError("solidboolean_b200 error %d: %s" % (code, msg))


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


class Context:
    """One device + one CUDA stream (sb_context)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _vp()
        _check(self.lib.sb_context_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self.lib.sb_context_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.sb_context_stream(self.h) or 0)

    def enable_timing(self, on=True):
        _check(self.lib.sb_context_enable_timing(self.h, 1 if on else 0))

    def reset_timing(self):
        _check(self.lib.sb_context_reset_timing(self.h))

    def timing(self):
        """-> ({stage: ms}, kernel launches) since the last reset (synchronises)."""
        ms = (C.c_float * len(STAGES))()
        n = C.c_uint64(0)
        _check(self.lib.sb_context_get_timing(self.h, ms, C.byref(n)))
        return {k: float(ms[i]) for i, k in enumerate(STAGES)}, int(n.value)

    def fp64_peak(self):
        """-> (GFLOP/s with separate DMUL+DADD, GFLOP/s with DFMA counted as 2)."""
        a, b = C.c_double(0), C.c_double(0)
        _check(self.lib.sb_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def classify_stats(self):
        r, c = C.c_uint64(0), C.c_uint64(0)
        _check(self.lib.sb_context_classify_stats(self.h, C.byref(r), C.byref(c)))
        return int(r.value), int(c.value)

    def mesh(self, xyz, tri, build=True) -> "Mesh":
        return Mesh(self, xyz, tri, build=build)

    def tri_tri_batch(self, tris18):
        """Raw predicate on explicit triangle pairs [n,18] -> ret, coplanar, seg[n,6]."""
        t = np.ascontiguousarray(tris18, dtype=np.float64).reshape(-1, 18)
        n = t.shape[0]
        ret = np.zeros(n, np.int32)
        cop = np.zeros(n, np.int32)
        seg = np.zeros((n, 6), np.float64)
        _check(self.lib.sb_tri_tri_batch(self.h, _ptr(t), n, _ptr(ret), _ptr(cop), _ptr(seg)))
        return ret, cop, seg


class Mesh:
    """Device-resident mesh + LBVH (sb_mesh); mirrors SolidMesh::prepare()."""

    def __init__(self, ctx: Context, xyz, tri, build=True):
        self.ctx = ctx
        self.lib = ctx.lib
        # keep the host arrays alive while an async upload may still read them
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
        h = _vp()
        fn = self.lib.sb_mesh_create if build else self.lib.sb_mesh_upload
        _check(fn(ctx.h, _ptr(self.xyz), self.xyz.shape[0], _ptr(self.tri), self.tri.shape[0], C.byref(h)))
        self.h = h

    @classmethod
    def from_pointers(cls, ctx: Context, xyz_ptr: int, nV: int, tri_ptr: int, nT: int, build=True, keep=None):
        """Create from raw host pointers (e.g. pinned torch tensors)."""
        self = cls.__new__(cls)
        self.ctx, self.lib = ctx, ctx.lib
        self.xyz = self.tri = None
        self._keep = keep
        h = _vp()
        fn = self.lib.sb_mesh_create if build else self.lib.sb_mesh_upload
        _check(fn(ctx.h, _vp(xyz_ptr), nV, _vp(tri_ptr), nT, C.byref(h)))
        self.h = h
        return self

    @classmethod
    def from_device(cls, ctx: Context, d_xyz: int, nV: int, d_tri: int, nT: int, keep=None):
        """sb_mesh_upload_device: geometry already on this device (no build)."""
        self = cls.__new__(cls)
        self.ctx, self.lib = ctx, ctx.lib
        self.xyz = self.tri = None
        self._keep = keep
        h = _vp()
        _check(self.lib.sb_mesh_upload_device(ctx.h, _vp(d_xyz), nV, _vp(d_tri), nT, C.byref(h)))
        self.h = h
        return self

    @classmethod
    def batch(cls, ctx: Context, jobs, lattice_pitch: float, build=True):
        """sb_batch_upload: `jobs` = sequence of (xyz, tri) small meshes (job-local indices), laid end to
        end into ONE batch mesh.  lattice_pitch >= 4 x the largest |coordinate| of both batches of a pair
        (`batch_pitch` computes it)."""
        xs = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3) for x, _ in jobs]
        ts = [np.ascontiguousarray(t, dtype=np.uint32).reshape(-1, 3) for _, t in jobs]
        return cls.batch_arrays(ctx, np.concatenate(xs) if xs else np.zeros((0, 3)),
                                np.concatenate([[0], np.cumsum([len(x) for x in xs])]),
                                np.concatenate(ts) if ts else np.zeros((0, 3), np.uint32),
                                np.concatenate([[0], np.cumsum([len(t) for t in ts])]), lattice_pitch, build)

    @classmethod
    def batch_arrays(cls, ctx: Context, xyz, vertex_start, tri, triangle_start, lattice_pitch: float, build=True):
        """The same from already concatenated arrays (+ their n_jobs + 1 start offsets)."""
        self = cls.__new__(cls)
        self.ctx, self.lib = ctx, ctx.lib
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
        self.vertex_start = np.ascontiguousarray(vertex_start, dtype=np.uint64)
        self.triangle_start = np.ascontiguousarray(triangle_start, dtype=np.uint64)
        h = _vp()
        _check(self.lib.sb_batch_upload(ctx.h, len(self.vertex_start) - 1, _ptr(self.xyz), _ptr(self.vertex_start),
                                        _ptr(self.tri), _ptr(self.triangle_start), float(lattice_pitch), C.byref(h)))
        self.h = h
        if build:
            self.build()
            ctx.synchronize()
        return self

    def update(self, xyz_ptr: int, tri_ptr: int, on_device=False):
        """sb_mesh_update: same-sized geometry into this mesh (raw pointers; 0 = keep)."""
        _check(self.lib.sb_mesh_update(self.h, _vp(xyz_ptr) if xyz_ptr else None, _vp(tri_ptr) if tri_ptr else None,
                                       1 if on_device else 0))

    def build(self):
        _check(self.lib.sb_mesh_build(self.h))

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.sb_mesh_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_triangles(self) -> int:
        return int(self.lib.sb_mesh_num_triangles(self.h))

    @property
    def num_vertices(self) -> int:
        return int(self.lib.sb_mesh_num_vertices(self.h))

    def normals(self):
        out = np.zeros((self.num_triangles, 3), np.float64)
        _check(self.lib.sb_mesh_normals(self.h, _ptr(out)))
        return out

    def triangle_boxes(self):
        out = np.zeros((self.num_triangles, 6), np.float64)
        _check(self.lib.sb_mesh_triangle_boxes(self.h, _ptr(out)))
        return out

    def bounds(self):
        out = np.zeros(6, np.float64)
        _check(self.lib.sb_mesh_bounds(self.h, _ptr(out)))
        return out

    def order(self):
        out = np.zeros(self.num_triangles, np.uint32)
        _check(self.lib.sb_mesh_order(self.h, _ptr(out)))
        return out

    def bvh(self):
        """-> dict(info, nodes [2*I] records, leaves [nTpad] records)."""
        info = _BvhInfo()
        _check(self.lib.sb_mesh_bvh_info(self.h, C.byref(info)))
        rec = np.dtype([("lo", np.float32, 3), ("hi", np.float32, 3), ("ref", np.int32), ("aux", np.int32)])
        nodes = np.zeros(2 * info.num_internal, rec)
        _check(self.lib.sb_mesh_bvh_nodes(self.h, _ptr(nodes)))
        npad = (self.num_triangles + 31) // 32 * 32
        leaves = np.zeros(npad, rec)
        cnt = _sz(0)
        _check(self.lib.sb_mesh_bvh_leaves(self.h, _ptr(leaves), C.byref(cnt)))
        assert cnt.value == npad
        return dict(cluster_size=info.cluster_size, num_clusters=info.num_clusters,
                    num_internal=info.num_internal, root=info.root, nodes=nodes, leaves=leaves)

    def grid_info(self):
        gi = _GridInfo()
        _check(self.lib.sb_mesh_grid_info(self.h, C.byref(gi)))
        return dict(nu=list(gi.nu), nv=list(gi.nv), total_cells=gi.total_cells, total_refs=gi.total_refs,
                    big=list(gi.big), mean_extent=list(gi.mean_extent))

    def intersect(self, other: "Mesh", flags=0, begin=None, end=None) -> "Isect":
        return Isect(self, other, flags, begin, end)

    def classify(self, pts, per_axis=True):
        """isPointInMesh majority vote for explicit points against THIS mesh.
        per_axis=False lets the library skip the third ray wherever the first two agree."""
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        q = p.shape[0]
        inside = np.zeros(q, np.uint8)
        axes = np.zeros((q, 3), np.uint8) if per_axis else None
        _check(self.lib.sb_classify(self.h, _ptr(p), q, _ptr(inside), _ptr(axes)))
        return inside, axes

    def classify_faces_against(self, target: "Mesh", per_axis=True, out=None):
        """Classify this mesh's face centroids against `target`.
        -> (inside[nT], per_axis[nT,3] or None); `out` = optional preallocated inside array."""
        n = self.num_triangles
        inside = np.zeros(n, np.uint8) if out is None else out
        axes = np.zeros((n, 3), np.uint8) if per_axis else None
        _check(self.lib.sb_classify_faces(self.h, target.h, _ptr(inside), _ptr(axes)))
        return inside, axes

    def uncut(self, cut_flags=None, vertex_offset=0, triangle_offset=0) -> "Uncut":
        """sb_mesh_uncut: addUnintersectedTriangles with an explicit per-face cut mask."""
        cf = None
        if cut_flags is not None:
            cf = np.ascontiguousarray(cut_flags, dtype=np.uint8)
            if cf.shape[0] != self.num_triangles:
                raise ValueError("cut_flags must hold one byte per triangle")
        h = _vp()
        _check(self.lib.sb_mesh_uncut(self.h, _ptr(cf), vertex_offset, triangle_offset, C.byref(h)))
        return Uncut(self, h)

    def classify_faces_device(self, target: "Mesh", d_inside_ptr: int, begin=0, end=None):
        end = self.num_triangles if end is None else end
        _check(self.lib.sb_classify_faces_device(self.h, target.h, begin, end, _vp(d_inside_ptr)))


class Shard:
    """One rank's share of a multi-GPU front end (sb_shard): `a`, `b` = uploaded meshes (build=False is enough)."""

    def __init__(self, a: Mesh, b: Mesh, rank: int, n_ranks: int):
        self.a, self.b, self.lib = a, b, a.lib
        h = _vp()
        _check(self.lib.sb_shard_create(a.h, b.h, rank, n_ranks, C.byref(h)))
        self.h = h

    def front_end(self, d_inside_a: int, d_inside_b: int, flags=0) -> "Isect":
        x = Isect.__new__(Isect)
        x.a, x.b, x.lib = self.a, self.b, self.lib
        h = _vp()
        _check(self.lib.sb_shard_front_end(self.h, flags, C.byref(h), _vp(d_inside_a), _vp(d_inside_b)))
        x.h = h
        nc, nh = _sz(0), _sz(0)
        _check(self.lib.sb_isect_counts(h, C.byref(nc), C.byref(nh)))
        x.num_candidates, x.num_hits = int(nc.value), int(nh.value)
        return x

    def info(self):
        sa, sb_, lo, hi, fb = _sz(0), _sz(0), C.c_double(0), C.c_double(0), C.c_uint64(0)
        _check(self.lib.sb_shard_info(self.h, C.byref(sa), C.byref(sb_), C.byref(lo), C.byref(hi), C.byref(fb)))
        return dict(selected_a=int(sa.value), selected_b=int(sb_.value), z_lo=float(lo.value), z_hi=float(hi.value),
                    fallbacks=int(fb.value))

    def close(self):
        if getattr(self, "h", None) and getattr(self.a.ctx, "h", None):
            self.lib.sb_shard_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """Several GPUs from one process (sb_comm_*): devices = list of CUDA device numbers, one rank each
    (the same device may appear twice)."""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        h = _vp()
        _check(self.lib.sb_comm_create(len(devices), devs, C.byref(h)))
        self.h = h
        self.nTA = self.nTB = 0
        self.num_candidates = self.num_hits = 0

    @property
    def size(self):
        return int(self.lib.sb_comm_size(self.h))

    def set_meshes(self, a, b):
        xa, ta = np.ascontiguousarray(a[0], np.float64), np.ascontiguousarray(a[1], np.uint32)
        xb, tb = np.ascontiguousarray(b[0], np.float64), np.ascontiguousarray(b[1], np.uint32)
        _check(self.lib.sb_comm_set_meshes(self.h, _ptr(xa), len(xa), _ptr(ta), len(ta), _ptr(xb), len(xb), _ptr(tb), len(tb)))
        self.nTA, self.nTB = len(ta), len(tb)

    def front_end(self, flags=0):
        """-> (insideA, insideB): per-face flags of both meshes; the counts land in num_candidates / num_hits."""
        ia, ib = np.zeros(self.nTA, np.uint8), np.zeros(self.nTB, np.uint8)
        nc, nh = _sz(0), _sz(0)
        _check(self.lib.sb_comm_front_end(self.h, flags, C.byref(nc), C.byref(nh), _ptr(ia), _ptr(ib)))
        self.num_candidates, self.num_hits = int(nc.value), int(nh.value)
        return ia, ib

    def hits(self):
        ab, seg = np.zeros((self.num_hits, 2), np.uint32), np.zeros((self.num_hits, 6), np.float64)
        _check(self.lib.sb_comm_hits(self.h, _ptr(ab), _ptr(seg)))
        return ab, seg

    def rank_info(self, r):
        i = _CommRankInfo()
        _check(self.lib.sb_comm_rank(self.h, r, C.byref(i)))
        return {k: getattr(i, k) for k, _ in _CommRankInfo._fields_}

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_comm_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def batch_pitch(*arrays) -> float:
    """A lattice pitch for sb_batch_upload: the power of two >= 4 x the largest |coordinate| of the given arrays."""
    m = max((float(np.abs(a).max()) for a in arrays if np.size(a)), default=1.0)
    p = 1.0
    while p < 4.0 * m:
        p *= 2.0
    while p * 0.5 >= 4.0 * m and p > 1e-300:
        p *= 0.5
    return p


class Isect:
    """Candidate pairs + intersecting pairs of two meshes (sb_isect)."""

    @classmethod
    def front_end_host(cls, a: Mesh, b: Mesh, inside_a_ptr: int, inside_b_ptr: int, hit_ab_ptr: int = 0, hit_seg_ptr: int = 0,
                       hit_capacity: int = 0, flags=0):
        """sb_front_end_host: the whole front end with HOST outputs (raw pointers, e.g. pinned torch tensors): per-face
        flags of both meshes and, when it fits hit_capacity, the hit list (pairs + segments)."""
        self = cls.__new__(cls)
        self.a, self.b, self.lib = a, b, a.lib
        h = _vp()
        _check(self.lib.sb_front_end_host(a.h, b.h, flags, C.byref(h), _vp(inside_a_ptr), _vp(inside_b_ptr),
                                          _vp(hit_ab_ptr) if hit_ab_ptr else None, _vp(hit_seg_ptr) if hit_seg_ptr else None,
                                          hit_capacity))
        self.h = h
        nc, nh = _sz(0), _sz(0)
        _check(self.lib.sb_isect_counts(h, C.byref(nc), C.byref(nh)))
        self.num_candidates, self.num_hits = int(nc.value), int(nh.value)
        return self

    @classmethod
    def front_end(cls, a: Mesh, b: Mesh, d_inside_a: int, d_inside_b: int, flags=0, a_range=None, b_range=None):
        """sb_front_end(_range): intersection + both classifications, overlapped."""
        self = cls.__new__(cls)
        self.a, self.b, self.lib = a, b, a.lib
        h = _vp()
        a0, a1 = a_range if a_range else (0, a.num_triangles)
        b0, b1 = b_range if b_range else (0, b.num_triangles)
        _check(self.lib.sb_front_end_range(a.h, b.h, a0, a1, b0, b1, flags, C.byref(h), _vp(d_inside_a), _vp(d_inside_b)))
        self.h = h
        nc, nh = _sz(0), _sz(0)
        _check(self.lib.sb_isect_counts(h, C.byref(nc), C.byref(nh)))
        self.num_candidates, self.num_hits = int(nc.value), int(nh.value)
        return self

    def __init__(self, a: Mesh, b: Mesh, flags=0, begin=None, end=None):
        self.a, self.b = a, b
        self.lib = a.lib
        h = _vp()
        if begin is None and end is None:
            _check(self.lib.sb_intersect(a.h, b.h, flags, C.byref(h)))
        else:
            _check(self.lib.sb_intersect_range(a.h, b.h, begin or 0, a.num_triangles if end is None else end,
                                               flags, C.byref(h)))
        self.h = h
        nc, nh = _sz(0), _sz(0)
        _check(self.lib.sb_isect_counts(h, C.byref(nc), C.byref(nh)))
        self.num_candidates, self.num_hits = int(nc.value), int(nh.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.a.ctx, "h", None):
            self.lib.sb_isect_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def candidates(self):
        ab = np.zeros((self.num_candidates, 2), np.uint32)
        code = np.zeros(self.num_candidates, np.uint8)
        _check(self.lib.sb_isect_candidates(self.h, _ptr(ab), _ptr(code)))
        return ab, code

    def hits(self):
        ab = np.zeros((self.num_hits, 2), np.uint32)
        seg = np.zeros((self.num_hits, 6), np.float64)
        _check(self.lib.sb_isect_hits(self.h, _ptr(ab), _ptr(seg)))
        return ab, seg

    def job_ranges(self):
        """Batch meshes: hits of job j = rows [r[j], r[j + 1]) of hits() (sb_batch_job_ranges)."""
        n = _sz(0)
        _check(self.lib.sb_batch_info(self.a.h, C.byref(n), None, None))
        out = np.zeros(n.value + 1, np.uint64)
        _check(self.lib.sb_batch_job_ranges(self.h, _ptr(out)))
        return out.astype(np.int64)

    def path_counts(self):
        """Predicate exit histogram: plane2 reject, plane1 reject, coplanar, interval reject, segment."""
        out = (C.c_uint64 * 5)()
        _check(self.lib.sb_isect_path_counts(self.h, out))
        return [int(v) for v in out]

    def face_flags(self):
        fa = np.zeros(self.a.num_triangles, np.uint8)
        fb = np.zeros(self.b.num_triangles, np.uint8)
        _check(self.lib.sb_isect_face_flags(self.h, _ptr(fa), _ptr(fb)))
        return fa, fb

    def contexts(self, which: int):
        """sb_isect_contexts: the per-triangle intersection contexts of mesh `which` (the pair-loop
        body of combine(), reference src/solidboolean.cpp:296-339), hits in ascending (a, b) order.
        -> dict(tri [c], point_start [c+1], points [p,3], edge_start [c+1], edges [e,2])"""
        h = _vp()
        _check(self.lib.sb_isect_contexts(self.h, which, C.byref(h)))
        try:
            nc, npts, ne = _sz(0), _sz(0), _sz(0)
            _check(self.lib.sb_cuts_counts(h, C.byref(nc), C.byref(npts), C.byref(ne)))
            tri = np.zeros(nc.value, np.uint32)
            ps = np.zeros(nc.value + 1, np.uint32)
            es = np.zeros(nc.value + 1, np.uint32)
            pts = np.zeros((npts.value, 3), np.float64)
            edges = np.zeros((ne.value, 2), np.uint32)
            _check(self.lib.sb_cuts_fetch(h, _ptr(tri), _ptr(ps), _ptr(pts), _ptr(es), _ptr(edges)))
        finally:
            self.lib.sb_cuts_destroy(h)
        return dict(tri=tri, point_start=ps, points=pts, edge_start=es, edges=edges)

    def hit_edges(self):
        """sb_isect_hit_edges -> tags [n_hit] uint8: bits 0-1 edge of the source point, bit 2 on B's triangle;
        bits 4-5 / 6 the same for the target point; bit 7 set."""
        tags = np.zeros(self.num_hits, np.uint8)
        if self.num_hits:
            _check(self.lib.sb_isect_hit_edges(self.h, _ptr(tags)))
        return tags

    def uncut(self, which: int, vertex_offset=0, triangle_offset=0) -> "Uncut":
        """sb_isect_uncut: the faces of mesh `which` the intersection left alone + their half-edge map."""
        h = _vp()
        _check(self.lib.sb_isect_uncut(self.h, which, vertex_offset, triangle_offset, C.byref(h)))
        return Uncut(self.a if which == 0 else self.b, h)

    def pack_device(self, d_record: int, cap: int):
        """sb_isect_pack_device: {nCand, nHit, hit pairs, segments} into a device buffer of 16 + 56 cap bytes."""
        _check(self.lib.sb_isect_pack_device(self.h, _vp(d_record), cap))

    def device_ptrs(self, candidates=True):
        """Device pointers of the results.  candidates=False leaves the candidate keys alone (asking
        for them orders the full candidate list on the device, which is otherwise done lazily)."""
        ck, ha, hs, fa, fb = _vp(), _vp(), _vp(), _vp(), _vp()
        bits = C.c_uint(0)
        _check(self.lib.sb_isect_device_ptrs(self.h, C.byref(ck) if candidates else None, C.byref(bits), C.byref(ha),
                                             C.byref(hs), C.byref(fa), C.byref(fb)))
        return dict(cand_keys=ck.value or 0, bits_b=bits.value, hit_ab=ha.value or 0, hit_seg=hs.value or 0,
                    flags_a=fa.value or 0, flags_b=fb.value or 0)


class Uncut:
    """Uncut triangles + half-edge map of one mesh (sb_uncut): what
    SolidBoolean::addUnintersectedTriangles leaves behind (reference src/solidboolean.cpp:250-286)."""

    def __init__(self, mesh: Mesh, h):
        self.mesh, self.lib, self.h = mesh, mesh.lib, h
        nt, nh, ok = _sz(0), _sz(0), C.c_int(0)
        _check(self.lib.sb_uncut_counts(h, C.byref(nt), C.byref(nh), C.byref(ok)))
        self.num_triangles, self.num_half_edges, self.ok = int(nt.value), int(nh.value), bool(ok.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.mesh.ctx, "h", None):
            self.lib.sb_uncut_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def triangles(self):
        """-> face [n] (original face id of each new triangle), tri3 [n,3] (shifted index triples)"""
        face = np.zeros(self.num_triangles, np.uint32)
        tri3 = np.zeros((self.num_triangles, 3), np.uint32)
        _check(self.lib.sb_uncut_triangles(self.h, _ptr(face), _ptr(tri3)))
        return face, tri3

    def half_edges(self):
        """-> keys [m] ascending ((first << 32) | second), owner [m] (new triangle index)"""
        keys = np.zeros(self.num_half_edges, np.uint64)
        owner = np.zeros(self.num_half_edges, np.uint32)
        _check(self.lib.sb_uncut_half_edges(self.h, _ptr(keys), _ptr(owner)))
        return keys, owner

    def adjacency(self):
        """-> adj [n,3]: triangle across edge k of new triangle j, or -1"""
        adj = np.full((self.num_triangles, 3), -1, np.int32)
        if self.num_triangles:
            _check(self.lib.sb_uncut_adjacency(self.h, _ptr(adj)))
        return adj

    def components(self):
        """-> label [n] (lowest new triangle index of each triangle's face group), group count"""
        label = np.zeros(self.num_triangles, np.uint32)
        n = _sz(0)
        _check(self.lib.sb_uncut_components(self.h, _ptr(label), C.byref(n)))
        return label, int(n.value)

    def face_groups(self, pieces, fences):
        """sb_uncut_face_groups: the flood over the uncut triangles and the retriangulated `pieces` [nP, 3] (result
        vertex ids), not crossing the `fences` [nF, 2] -> label_uncut [n], label_piece [nP] (lowest node of the group;
        nodes = uncut triangles 0..n-1, then pieces), group count"""
        pc = np.ascontiguousarray(pieces, np.uint32).reshape(-1, 3)
        fc = np.ascontiguousarray(fences, np.uint32).reshape(-1, 2)
        lu, lp = np.zeros(self.num_triangles, np.uint32), np.zeros(len(pc), np.uint32)
        n = _sz(0)
        _check(self.lib.sb_uncut_face_groups(self.h, _ptr(pc) if len(pc) else None, len(pc), _ptr(fc) if len(fc) else None, len(fc),
                                             _ptr(lu) if len(lu) else None, _ptr(lp) if len(lp) else None, C.byref(n)))
        return lu, lp, int(n.value)

    def device_ptrs(self):
        p = [_vp() for _ in range(5)]
        _check(self.lib.sb_uncut_device_ptrs(self.h, *[C.byref(x) for x in p]))
        return dict(zip(("face", "tri3", "keys", "owner", "adj"), [x.value or 0 for x in p]))
