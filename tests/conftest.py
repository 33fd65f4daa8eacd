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
