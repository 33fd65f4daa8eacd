"""Per-source-line instruction / stall totals of one kernel launch in an .ncu-rep
(needs -lineinfo and --import-source on):
    python scripts/ncu_lines.py rep [launch-skip] [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur = None; fname = ""
inst = collections.Counter(); samples = collections.Counter(); thr = collections.Counter(); text = {}
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iI = r.index("Instructions Executed"); iS = r.index("# Samples"); iT = r.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0]:
        cur = (fname, int(r[0])); text[cur] = r[1]; continue
    if cur and r[iI].isdigit():
        inst[cur] += int(r[iI]); samples[cur] += int(r[iS]) if r[iS].isdigit() else 0; thr[cur] += int(r[iT])
tot = sum(inst.values()); ts = sum(samples.values())
print("total warp instructions %d, samples %d" % (tot, ts))
for k, v in inst.most_common(top):
    print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%d  %s" % (100 * v / tot, 100 * samples[k] / max(ts, 1), thr[k] / max(v, 1), k[0], k[1], text[k].strip()[:90]))
if len(sys.argv) > 4:  # file-level totals
    byfile = collections.Counter(); sf = collections.Counter()
    for k, v in inst.items(): byfile[k[0]] += v; sf[k[0]] += samples[k]
    for k, v in byfile.most_common(): print("%5.1f%% inst %5.1f%% smp  %s" % (100 * v / tot, 100 * sf[k] / max(ts, 1), k))
    rng = [int(x) for x in sys.argv[4].split(",")]
    main = [k for k in inst if k[0] == "sb_classify.cu"]
    for a, b in zip(rng[:-1], rng[1:]):
        v = sum(inst[k] for k in main if a <= k[1] < b); s = sum(samples[k] for k in main if a <= k[1] < b)
        print("  sb_classify.cu:%d-%d  %5.1f%% inst %5.1f%% smp" % (a, b - 1, 100 * v / tot, 100 * s / max(ts, 1)))
