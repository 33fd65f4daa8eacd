"""Whole SolidBoolean::combine() + fetch* through the C++ mirror of the reference API
(solidboolean_b200/host) on one GPU (dev / measurement tool):
    python scripts/combine_times.py [c2|c3] [reps]
Prints the reference's own stage time-points (src/solidboolean.h:45-58) per repetition and checks
the three results as solids (closed manifolds, inclusion-exclusion of the volumes)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from solidboolean_b200 import meshgen
from test_host_cpp import _bind as bind, run_boolean, MOCK_SO   # the ctypes view of sbh_capi.cpp

args = [x for x in sys.argv[1:] if not x.startswith("--")]
cfg = args[0] if args else "c3"
reps = int(args[1]) if len(args) > 1 else 2
a, b = {"c2": meshgen.config_c2, "c3": meshgen.config_c3}[cfg]()
# --mock: the oracle-backed CPU stand-in of the C ABI (tests/hostsim), to time the HOST stages without a GPU
so = MOCK_SO if "--mock" in sys.argv else os.path.join("solidboolean_b200", "lib", "libsolidboolean_host.so")
lib = bind(C.CDLL(os.path.abspath(so)))
names = ("search", "process", "addUnintersected", "reTriangulate", "buildPolygons", "buildFaceGroups", "decideGroupSide")
for it in range(reps):
    t0 = time.perf_counter()
    res = run_boolean(lib, a, b)
    wall = (time.perf_counter() - t0) * 1e3
    rec = dict(it=it, ok=res["ok"], wall_ms_incl_prepare_and_fetch=round(wall, 1), P=res["P"], H=res["H"],
               stage_ms={k: round(float(v), 2) for k, v in zip(names, res["stage_ms"])},
               combine_ms=round(float(sum(res["stage_ms"])), 2))
    if not res["ok"]:
        rec["log"] = res["log"][-300:]
    print(json.dumps(rec), flush=True)
if res["ok"]:
    v = res["vertices"]
    va, vb = meshgen.signed_volume(*a), meshgen.signed_volume(*b)
    vol = {k: meshgen.signed_volume(v, res[k]) for k in ("union", "diff", "intersect")}
    print(json.dumps(dict(config=cfg, triangles={k: int(len(res[k])) for k in vol}, volumes=vol,
                          closed={k: bool(meshgen.is_closed_manifold(res[k])) for k in vol},
                          inclusion_exclusion_error=abs(vol["union"] + vol["intersect"] - va - vb),
                          diff_error=abs(vol["diff"] + vol["intersect"] - va))), flush=True)
