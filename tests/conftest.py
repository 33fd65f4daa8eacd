import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["simple-ring", "cube-sphere", "complex", "addax-and-meerkat"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def golden_json():
    import json
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_cases():
    return np.load(os.path.join(GOLDEN, "cases.npz"))


@pytest.fixture(scope="session")
def golden_synthetic():
    return np.load(os.path.join(GOLDEN, "synthetic.npz"))


@pytest.fixture(scope="session")
def golden_kat():
    return np.load(os.path.join(GOLDEN, "tritri_kat.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle.get()


def load_case(golden_cases, case):
    """-> (meshA, meshB, outputs dict) for one bundled reference test case."""
    key = case.replace("-", "_")
    a = (golden_cases[key + "__xyz_a"].astype(np.float64), golden_cases[key + "__tri_a"])
    b = (golden_cases[key + "__xyz_b"].astype(np.float64), golden_cases[key + "__tri_b"])
    out = {k: golden_cases[key + "__" + k] for k in
           ("pairs", "ret", "coplanar", "hit", "seg_hits", "inside_a", "per_axis_a", "inside_b", "per_axis_b")}
    return a, b, out


def synthetic_specs():
    from solidboolean_b200 import meshgen
    return {
        "ico3_offset": lambda: (meshgen.icosphere(3), meshgen.icosphere(3, center=(0.71, 0.13, 0.07))),
        "ico4_f32": lambda: (meshgen.icosphere(4, round_to_float=True),
                             meshgen.icosphere(4, center=(0.71, 0.13, 0.07), round_to_float=True)),
        "ico4_torus": lambda: (meshgen.icosphere(4), meshgen.torus(48, 24, center=(0.013, 0.007, 0.011))),
        "ico3_coincident": lambda: (meshgen.icosphere(3), meshgen.icosphere(3)),
        "slab_cross": lambda: (meshgen.slab(12, 1.0, 0.2),
                               meshgen.slab(12, 1.0, 0.2, center=(0.3, 0.1, 0.05), tilt=0.3)),
        "ico5_near": lambda: (meshgen.icosphere(5), meshgen.icosphere(5, center=(1.5e-3, 0.6e-3, 0.3e-3))),
    }


def load_synthetic(golden_synthetic, name):
    a, b = synthetic_specs()[name]()
    out = {k: golden_synthetic[name + "__" + k] for k in
           ("pairs", "ret", "coplanar", "hit", "seg_hits", "inside_a", "per_axis_a", "inside_b", "per_axis_b")}
    return a, b, out


def uncut_inputs():
    """name -> (mesh a, mesh b, cut flags a, cut flags b): the bundled cases and the synthetic
    pairs with the faces their (golden) hit pairs touch, plus crafted meshes."""
    from solidboolean_b200 import meshgen
    cases = np.load(os.path.join(GOLDEN, "cases.npz"))
    syn = np.load(os.path.join(GOLDEN, "synthetic.npz"))
    out = {}

    def flags(n, ids):
        f = np.zeros(n, np.uint8)
        f[ids] = 1
        return f

    for case in CASES:
        k = case.replace("-", "_")
        a = (cases[k + "__xyz_a"].astype(np.float64), cases[k + "__tri_a"])
        b = (cases[k + "__xyz_b"].astype(np.float64), cases[k + "__tri_b"])
        hp = cases[k + "__pairs"][cases[k + "__hit"].astype(bool)]
        out[case] = (a, b, flags(len(a[1]), hp[:, 0]), flags(len(b[1]), hp[:, 1]))
    for name, make in synthetic_specs().items():
        a, b = make()
        hp = syn[name + "__pairs"][syn[name + "__hit"].astype(bool)]
        out[name] = (a, b, flags(len(a[1]), hp[:, 0]), flags(len(b[1]), hp[:, 1]))
    # crafted: a triangle present twice (the reference stops at the copy), an open sheet, a
    # flipped triangle (its half-edges repeat its neighbours'), nothing cut, everything cut
    ico = meshgen.icosphere(2)
    tor = meshgen.torus(16, 8)
    dup = (ico[0], np.concatenate([ico[1][:40], ico[1][7:8], ico[1][40:]]))
    out["repeat_first"] = (dup, tor, None, None)
    out["repeat_second"] = (tor, dup, None, flags(len(dup[1]), [3, 4, 5]))
    flipped = ico[1].copy()
    flipped[100] = flipped[100][::-1]
    out["flipped_face"] = ((ico[0], flipped), tor, flags(len(flipped), [0, 1]), None)
    sheet = meshgen.slab(6, 1.0, 0.2)
    out["open_sheet"] = ((sheet[0], sheet[1][: len(sheet[1]) // 3]), tor, None, flags(len(tor[1]), np.arange(0, len(tor[1]), 5)))
    out["all_cut"] = (ico, tor, np.ones(len(ico[1]), np.uint8), np.ones(len(tor[1]), np.uint8))
    return out


def check_contexts_against(o, cap_get, side):
    """o: Oracle.cut_contexts-style dict (tri ascending, CSR points / edges) of ALL contexts;
    cap_get(key): the reference's captured arrays of one side ("tri", "point_start", ...), possibly a
    subset (combine() stopped early).  Points bit for bit, edges exactly."""
    tri = cap_get("tri")
    idx = np.searchsorted(o["tri"], tri)
    assert np.all(idx < max(len(o["tri"]), 1)) or len(tri) == 0
    assert np.array_equal(o["tri"][idx], tri), side + ": context triangles differ"
    ps, pts, es, ed = cap_get("point_start"), cap_get("points"), cap_get("edge_start"), cap_get("edges")
    for j, i in enumerate(idx):
        mine = o["points"][o["point_start"][i]:o["point_start"][i + 1]]
        assert mine.tobytes() == pts[ps[j]:ps[j + 1]].tobytes(), (side, int(tri[j]), "points")
        assert np.array_equal(o["edges"][o["edge_start"][i]:o["edge_start"][i + 1]], ed[es[j]:es[j + 1]]), (side, int(tri[j]), "edges")
