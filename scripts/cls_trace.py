"""Dev tool: summarise the per-CTA timeline written with SB_CLASSIFY_TRACE=<file> (last launches in the file)."""
import sys, numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
i = 0; launches = []
while i < len(raw):
    assert raw[i] == np.uint64(0xffffffffffffffff)
    nb = int(raw[i + 1]); launches.append(raw[i + 4:i + 4 + 4 * nb].reshape(nb, 4)); i += 4 + 4 * nb
for L in launches[-2:]:
    sm, t0, t1, ent = L[:, 0].astype(int), L[:, 1].astype(np.int64), L[:, 2].astype(np.int64), L[:, 3].astype(np.int64)
    T0 = t0.min(); t0 -= T0; t1 -= T0
    dur = t1 - t0; total = t1.max()
    print("launch: %d CTAs, kernel span %.1f us, CTA duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (
        len(L), total / 1e3, dur.mean() / 1e3, np.percentile(dur, 50) / 1e3, np.percentile(dur, 90) / 1e3, np.percentile(dur, 99) / 1e3, dur.max() / 1e3))
    print("  entries/CTA: mean %.0f max %d; CTAs with 0 entries: %d" % (ent.mean(), ent.max(), (ent == 0).sum()))
    edges = np.linspace(0, total, 11)
    for a, b in zip(edges[:-1], edges[1:]):
        busy = len(set(sm[(t0 < b) & (t1 > a)]))
        conc = (np.minimum(t1, b) - np.maximum(t0, a)).clip(0).sum() / (b - a)
        e = ent[(t0 >= a) & (t0 < b)].sum()
        print("  %6.1f-%6.1f us: SMs busy %3d, CTAs in flight %.0f, entries started %d" % (a / 1e3, b / 1e3, busy, conc, e))
    k = np.argsort(-dur)[:5]
    print("  slowest CTAs (id, us, entries, start us):", [(int(x), round(dur[x] / 1e3, 1), int(ent[x]), round(t0[x] / 1e3, 1)) for x in k])
    print("  duration vs entries correlation: %.2f; us per 1000 entries (CTAs > 2000 entries): %.1f" % (
        np.corrcoef(dur, ent)[0, 1], (1e0 * dur[ent > 2000] / ent[ent > 2000]).mean() if (ent > 2000).any() else 0))
