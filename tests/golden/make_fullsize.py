#!/usr/bin/env python
"""Full-size pins from the UNMODIFIED reference (oracle/_ref/libsbref.so), generated in the build container:

    python tests/golden/make_fullsize.py [c3] [c4k8]

For BASELINE configs 3 (icosphere k=8 vs torus 1024x512) and 4 at its stated size (near-coincident icospheres
k=8, ~8.2 M candidate pairs) the reference's own prepare() / searchPotentialIntersectedPairs / predicate loop /
isPointInMesh run over ALL faces (the classification spread over the host threads), and what they return is
pinned as counts + FNV-1a-64 hashes in tests/golden/fullsize.json:
    pairs        sorted candidate pairs (uint64 a, b)
    codes        per sorted pair: ret | coplanar << 1 (uint8)
    hits / seg   accepted pairs and their segments (6 doubles each, bit patterns)
    inside_a/b   per-face inside flags (majority of the three rays), original face order
    per_axis_a/b per-face, per-axis flags (3 bytes per face)
tests/test_gpu_parity.py compares the CUDA path's outputs with these hashes at full size.
"""
import json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Ref, fnv1a64  # noqa: E402
from solidboolean_b200 import meshgen  # noqa: E402
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "fullsize.json")


def h(a):
    return "%016x" % fnv1a64(np.ascontiguousarray(a).tobytes())


def run(name, a, b, threads):
    R = Ref.get()
    t0 = time.time()
    out = [None, None]
    ts = [threading.Thread(target=lambda i=i, m=m: out.__setitem__(i, R.mesh(*m))) for i, m in enumerate((a, b))]
    [t.start() for t in ts]; [t.join() for t in ts]
    ma, mb = out
    op = R.op(ma, mb)
    pr = op.search()
    prs = pr[np.lexsort((pr[:, 1], pr[:, 0]))]
    ret, cop, hit, seg, _ = op.predicate(prs)
    print(name, "pairs", len(prs), "hits", int(hit.sum()), "%.1fs" % (time.time() - t0), flush=True)
    res = {}
    for tgt, cen, key in ((1, ma.centroids(), "a"), (0, mb.centroids(), "b")):
        chunks = np.array_split(np.arange(len(cen)), threads * 4)
        ins = np.zeros(len(cen), np.uint8); per = np.zeros((len(cen), 3), np.uint8)

        def work(ix):
            i, p, _ = op.classify(tgt, cen[ix])
            ins[ix] = i; per[ix] = p
        ws = [threading.Thread(target=work, args=(ix,)) for ix in chunks]
        for i in range(0, len(ws), threads):
            [w.start() for w in ws[i:i + threads]]; [w.join() for w in ws[i:i + threads]]
        res["inside_" + key] = ins; res["per_axis_" + key] = per
        print(name, "classified", key, int(ins.sum()), "%.1fs" % (time.time() - t0), flush=True)
    hb = hit.astype(bool)
    d = dict(tris_a=len(a[1]), tris_b=len(b[1]), n_pairs=int(len(prs)), n_hits=int(hb.sum()),
             pairs=h(prs.astype("<u8")), codes=h((ret | (cop << 1)).astype(np.uint8)),
             hits=h(prs[hb].astype("<u8")), seg=h(seg[hb].astype("<f8")),
             inside_a_count=int(res["inside_a"].sum()), inside_b_count=int(res["inside_b"].sum()),
             inside_a=h(res["inside_a"]), inside_b=h(res["inside_b"]),
             per_axis_a=h(res["per_axis_a"]), per_axis_b=h(res["per_axis_b"]))
    op.close(); ma.close(); mb.close()
    return d


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4k8"]
    gens = {"c3": meshgen.config_c3, "c4k8": lambda: meshgen.config_c4(k=8)}
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    threads = len(os.sched_getaffinity(0))
    for w in which:
        a, b = gens[w]()
        data[w] = run(w, a, b, threads)
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
        print(w, data[w], flush=True)
