"""The C++ mirror of the reference API (solidboolean_b200/host): SolidMesh::prepare,
SolidBoolean::combine, fetchUnion/Diff/Intersect.

CPU tests run the HOST logic (cutting, welding, face groups, assembly) against a
test-only oracle-backed stand-in for the C ABI (tests/hostsim/sb_mock_abi.cpp);
the -m gpu tests run the real thing (libsolidboolean_host.so -> CUDA library).
Results are compared with the reference as SOLIDS (closed, same volume, same
front-end counts): the reference's own triangulation depends on its tree's pair
order and on unordered_map iteration (SURVEY hard part 4)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES, load_case, load_synthetic
from solidboolean_b200 import meshgen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(ROOT, "solidboolean_b200", "host")
MOCK_SO = os.path.join(HERE, "hostsim", "libsbhost_mock.so")
REAL_SO = os.path.join(ROOT, "solidboolean_b200", "lib", "libsolidboolean_host.so")


def _bind(lib):
    vp = C.c_void_p
    lib.sbh_boolean.restype = vp
    lib.sbh_boolean.argtypes = [vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t]
    lib.sbh_ok.argtypes = [vp]
    lib.sbh_log.argtypes = [vp]
    lib.sbh_log.restype = C.c_char_p
    for name in ("sbh_candidates", "sbh_hits", "sbh_vertex_count"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = C.c_size_t
    lib.sbh_vertices.argtypes = [vp, vp]
    lib.sbh_triangle_count.argtypes = [vp, C.c_int]
    lib.sbh_triangle_count.restype = C.c_size_t
    lib.sbh_triangles.argtypes = [vp, C.c_int, vp]
    lib.sbh_stage_ms.argtypes = [vp, vp]
    lib.sbh_free.argtypes = [vp]
    return lib


@pytest.fixture(scope="module")
def mock_lib():
    from oracle import ORACLE_SO, build
    if not os.path.exists(ORACLE_SO):
        build()
    srcs = [os.path.join(HOST, f) for f in ("solidmesh.cpp", "solidboolean.cpp", "retriangulator.cpp", "sbh_capi.cpp")]
    srcs.append(os.path.join(HERE, "hostsim", "sb_mock_abi.cpp"))
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
    if not os.path.exists(MOCK_SO) or any(os.path.getmtime(d) > os.path.getmtime(MOCK_SO) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-I" + HOST, "-I" + os.path.join(ROOT, "include"),
                        "-o", MOCK_SO] + srcs + [ORACLE_SO, "-Wl,-rpath," + os.path.dirname(ORACLE_SO)],
                       check=True, capture_output=True)
    return _bind(C.CDLL(MOCK_SO))


def run_boolean(lib, a, b):
    xa, ta = np.ascontiguousarray(a[0], np.float64), np.ascontiguousarray(a[1], np.uint32)
    xb, tb = np.ascontiguousarray(b[0], np.float64), np.ascontiguousarray(b[1], np.uint32)
    vp = C.c_void_p
    h = lib.sbh_boolean(xa.ctypes.data_as(vp), len(xa), ta.ctypes.data_as(vp), len(ta),
                        xb.ctypes.data_as(vp), len(xb), tb.ctypes.data_as(vp), len(tb))
    res = {"ok": bool(lib.sbh_ok(h)), "log": lib.sbh_log(h).decode(), "P": lib.sbh_candidates(h), "H": lib.sbh_hits(h)}
    if res["ok"]:
        v = np.zeros((lib.sbh_vertex_count(h), 3), np.float64)
        lib.sbh_vertices(h, v.ctypes.data_as(vp))
        res["vertices"] = v
        for which, name in enumerate(("union", "diff", "intersect")):
            t = np.zeros((lib.sbh_triangle_count(h, which), 3), np.uint32)
            if len(t):
                lib.sbh_triangles(h, which, t.ctypes.data_as(vp))
            res[name] = t
    st = np.zeros(7)
    lib.sbh_stage_ms(h, st.ctypes.data_as(vp))
    res["stage_ms"] = st
    lib.sbh_free(h)
    return res


def check_solid(res, meta, a, b, vol_rtol=1e-6):
    assert meta.get("combine_ok", True) is True          # the reference's own combine() returned true on this input ...
    assert res["ok"], res["log"]                         # ... and so does the mirror's
    assert (res["P"], res["H"]) == (meta["P"], meta["H"])
    v = res["vertices"]
    # result vertex array = A's vertices, then B's, then welded new points
    assert np.array_equal(v[:len(a[0])], a[0]) and np.array_equal(v[len(a[0]):len(a[0]) + len(b[0])], b[0])
    assert len(v) == meta["result_vertices"]
    for name in ("union", "diff", "intersect"):
        t = res[name]
        assert len(t) > 0
        assert meshgen.is_closed_manifold(t), name + " is not a closed manifold"
        vol = meshgen.signed_volume(v, t)
        ref = meta["volume_" + name]
        assert abs(vol - ref) <= vol_rtol * max(abs(ref), 1e-12), (name, vol, ref)
    va, vb = meshgen.signed_volume(*a), meshgen.signed_volume(*b)
    vu = meshgen.signed_volume(v, res["union"])
    vi = meshgen.signed_volume(v, res["intersect"])
    vd = meshgen.signed_volume(v, res["diff"])
    assert abs(vu + vi - (va + vb)) <= 1e-9 * (abs(va) + abs(vb))   # inclusion-exclusion
    assert abs(vd + vi - va) <= 1e-9 * (abs(va) + abs(vb))


@pytest.mark.parametrize("case", CASES)
def test_host_logic_bundled_cases_cpu(mock_lib, golden_cases, golden_json, case):
    a, b, _ = load_case(golden_cases, case)
    res = run_boolean(mock_lib, a, b)
    check_solid(res, golden_json["cases"][case], a, b)


@pytest.mark.parametrize("name", ["ico3_offset", "ico4_f32", "ico4_torus"])
def test_host_logic_synthetic_cpu(mock_lib, golden_synthetic, golden_json, name):
    a, b, _ = load_synthetic(golden_synthetic, name)
    res = run_boolean(mock_lib, a, b)
    check_solid(res, golden_json["synthetic"][name], a, b)


@pytest.mark.parametrize("name", ["ico3_coincident", "ico5_near", "slab_cross"])
def test_host_logic_where_the_reference_gives_up_cpu(mock_lib, golden_synthetic, golden_json, name):
    """Coincident / near-coincident / welded inputs: the reference's own combine() returns false on
    these fixtures (golden.json: combine_ok false -- exact collinearity test, src/vector2.h:212-215).
    The mirror may get further (its attach test has a tolerance, branching curves are walked greedily
    as src/retriangulator.cpp:48-95 does), but whatever it returns must be consistent: either false with
    a message on the log, or three results that satisfy inclusion-exclusion."""
    meta = golden_json["synthetic"][name]
    assert meta["combine_ok"] is False
    a, b, _ = load_synthetic(golden_synthetic, name)
    res = run_boolean(mock_lib, a, b)
    assert (res["P"], res["H"]) == (meta["P"], meta["H"])         # the front end is the same either way
    if not res["ok"]:
        assert res["log"].strip() != ""
        return
    v = res["vertices"]
    va, vb = meshgen.signed_volume(*a), meshgen.signed_volume(*b)
    vu = meshgen.signed_volume(v, res["union"]); vi = meshgen.signed_volume(v, res["intersect"])
    assert abs(vu + vi - (va + vb)) <= 1e-6 * (abs(va) + abs(vb))


def test_host_logic_disjoint_and_nested_cpu(mock_lib):
    a = meshgen.icosphere(2)
    far = meshgen.icosphere(2, center=(5, 0, 0))
    res = run_boolean(mock_lib, a, far)
    assert res["ok"] and res["H"] == 0
    assert len(res["union"]) == len(a[1]) + len(far[1]) and len(res["intersect"]) == 0
    assert len(res["diff"]) == len(a[1])
    inner = meshgen.icosphere(2, radius=0.4)
    res = run_boolean(mock_lib, a, inner)
    assert res["ok"] and res["H"] == 0
    assert len(res["union"]) == len(a[1]) and len(res["intersect"]) == len(inner[1])
    assert len(res["diff"]) == len(a[1]) + len(inner[1])           # shell: outer + reversed inner
    v = res["vertices"]
    assert abs(meshgen.signed_volume(v, res["diff"]) -
               (meshgen.signed_volume(*a) - meshgen.signed_volume(*inner))) < 1e-12


def test_earclip_triangulates_polygon_with_hole(mock_lib):
    # exercised through a triangle pierced by a small prism (closed loop inside one face)
    big = (np.array([[-2, -2, 0], [2, -2, 0], [0, 2, 0], [0, 0, -1.5]], np.float64),
           np.array([[0, 1, 2], [0, 3, 1], [1, 3, 2], [2, 3, 0]], np.uint32))
    assert meshgen.signed_volume(*big) > 0 and meshgen.is_closed_manifold(big[1])
    prism = meshgen.icosphere(2, radius=0.3, center=(0.013, -0.3, 0.047))
    res = run_boolean(mock_lib, big, prism)
    assert res["ok"], res["log"]
    v = res["vertices"]
    for name in ("union", "diff", "intersect"):
        assert meshgen.is_closed_manifold(res[name]), name
    va, vb = meshgen.signed_volume(*big), meshgen.signed_volume(*prism)
    vu = meshgen.signed_volume(v, res["union"]); vi = meshgen.signed_volume(v, res["intersect"])
    assert abs(vu + vi - (va + vb)) < 1e-9


# ---------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def real_lib():
    if not os.path.exists(REAL_SO):
        subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return _bind(C.CDLL(REAL_SO))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_drop_in_classes_bundled_cases_gpu(real_lib, golden_cases, golden_json, case):
    a, b, _ = load_case(golden_cases, case)
    res = run_boolean(real_lib, a, b)
    check_solid(res, golden_json["cases"][case], a, b)


@pytest.mark.gpu
def test_drop_in_classes_finer_than_the_reference_can_cut_gpu(real_lib):
    """Config C2 (81,920 + 81,920): the reference's own combine() fails here
    ("Attach point to triangle edge failed"); ours completes."""
    a, b = meshgen.config_c2()
    res = run_boolean(real_lib, a, b)
    assert res["ok"], res["log"][-400:]
    assert (res["P"], res["H"]) == (7754, 1386)
    v = res["vertices"]
    for name in ("union", "diff", "intersect"):
        assert meshgen.is_closed_manifold(res[name]), name
    va, vb = meshgen.signed_volume(*a), meshgen.signed_volume(*b)
    vu = meshgen.signed_volume(v, res["union"]); vi = meshgen.signed_volume(v, res["intersect"])
    assert abs(vu + vi - (va + vb)) < 1e-8


@pytest.mark.gpu
@pytest.mark.slow
def test_drop_in_classes_config_c3_gpu(real_lib):
    """Config C3 (1,310,720 + 1,048,576): the whole combine() + all three fetch* through the
    reference's class API -- front end, uncut triangles, half-edge map and their face groups
    on the GPU, retriangulation of the cut faces on the host."""
    a, b = meshgen.config_c3()
    res = run_boolean(real_lib, a, b)
    assert res["ok"], res["log"][-400:]
    assert (res["P"], res["H"]) == (36125, 9606)
    v = res["vertices"]
    for name in ("union", "diff", "intersect"):
        # Known limit: where the curve passes through a mesh vertex the cut
        # faces around it get a new point ~4e-7 away from that vertex while the face it only touches
        # keeps the original vertex; the reference's algorithm never welds the two, so a few dozen of
        # ~4 M half-edges have no partner.  Same count with SB_HOST_FLOOD=legacy.
        t = res[name].astype(np.int64)
        he = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
        keys = he[:, 0] << 32 | he[:, 1]
        assert len(np.unique(keys)) == len(keys), name + ": repeated half-edge"
        unmatched = int((~np.isin(he[:, 1] << 32 | he[:, 0], keys)).sum())
        assert unmatched <= 64, (name, unmatched)
    va, vb = meshgen.signed_volume(*a), meshgen.signed_volume(*b)
    vu = meshgen.signed_volume(v, res["union"]); vi = meshgen.signed_volume(v, res["intersect"])
    vd = meshgen.signed_volume(v, res["diff"])
    assert abs(vu + vi - (va + vb)) < 1e-6 and abs(vd + vi - va) < 1e-6
    assert 0 < vi < min(va, vb)


def _write_obj(path, xyz, tri):
    with open(path, "w") as f:
        for v in xyz:
            f.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))   # 9 digits: the float the reference's loader parsed
        for t in tri:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def _read_obj(path):
    v, t = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                v.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                t.append([int(x.split("/")[0]) - 1 for x in p[1:4]])
    return np.array(v, np.float64), np.array(t, np.int64)


@pytest.mark.gpu
def test_reference_main_cpp_unchanged_runs_on_the_gpu(tmp_path, golden_cases, golden_json):
    """BASELINE config 1: the reference's own test/main.cpp, compiled UNCHANGED against the drop-in
    headers (solidboolean_b200/host/Makefile `refmain`, built where the reference tree is present;
    the binary travels with the snapshot), executed on the GPU box.  It loads
    ../../cases/addax-and-meerkat/{a,b}.obj, runs prepare() x2 + combine() + the three fetch* and
    writes debug-merged-result.obj (union, diff at x-2, intersect at x+2).  The OBJ pair is rewritten
    from the committed fixture (the reference tree does not exist on the GPU box); the result is
    compared with the reference's own run as a SOLID: summed volume of the three results, and
    every result closed."""
    exe = os.path.join(HOST, "test-solidboolean")
    if not os.path.exists(exe):
        pytest.skip("test-solidboolean not built (needs the reference tree at build time)")
    a, b, _ = load_case(golden_cases, "addax-and-meerkat")
    case_dir = tmp_path / "cases" / "addax-and-meerkat"
    run_dir = tmp_path / "build" / "run"
    case_dir.mkdir(parents=True); run_dir.mkdir(parents=True)
    _write_obj(case_dir / "a.obj", *a)
    _write_obj(case_dir / "b.obj", *b)
    r = subprocess.run([exe], cwd=run_dir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    assert "Duration:" in r.stdout and "searchPotentialIntersectedPairs:" in r.stdout
    v, t = _read_obj(run_dir / "debug-merged-result.obj")
    meta = golden_json["cases"]["addax-and-meerkat"]
    assert len(v) == 3 * meta["result_vertices"]          # the result vertex array, once per operation
    # split the merged file back into the three results by their vertex blocks
    nv = meta["result_vertices"]
    block = t[:, 0] // nv
    total = 0.0
    for k, name in enumerate(("union", "diff", "intersect")):
        tk = t[block == k] - k * nv
        assert len(tk) > 0 and np.all((t[block == k] // nv) == k)
        assert meshgen.is_closed_manifold(tk.astype(np.uint32)), name
        vol = meshgen.signed_volume(v[k * nv:(k + 1) * nv], tk)
        ref = meta["volume_" + name]
        assert abs(vol - ref) <= 1e-3 * abs(ref), (name, vol, ref)   # the OBJ text is written with %f: 6 decimals
        total += vol
    assert abs(total - (meta["volume_union"] + meta["volume_diff"] + meta["volume_intersect"])) <= 1e-3 * abs(total)


@pytest.mark.gpu
def test_concurrent_booleans_from_several_threads_gpu(real_lib, golden_cases, golden_json):
    """The reference allows concurrent SolidBoolean objects from several threads ("Threading").
    Here they share the process's context, whose entry points lock it: four threads x two rounds of
    different bundled cases must each give the result of a run on its own."""
    import threading
    cases = ["simple-ring", "cube-sphere", "complex", "addax-and-meerkat"]
    inputs = {c: load_case(golden_cases, c)[:2] for c in cases}
    out, errs = {}, []

    def work(k):
        try:
            for rnd in range(2):
                c = cases[(k + rnd) % len(cases)]
                out[(k, rnd)] = (c, run_boolean(real_lib, *inputs[c]))
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert len(out) == 8
    for (k, rnd), (c, res) in out.items():
        a, b = inputs[c]
        check_solid(res, golden_json["cases"][c], a, b)
