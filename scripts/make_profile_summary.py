"""Write profiles/<round>_ncu_summary.md + profiles/traffic.json from the ncu artefacts in gpurun_out/:
    python scripts/make_profile_summary.py launches.csv step_full.ncu-rep [r02]"""
import collections, csv, json, subprocess, sys
launch_csv, rep = sys.argv[1], sys.argv[2]
RND = sys.argv[3] if len(sys.argv) > 3 else "r02"
out_md, out_json = "profiles/%s_ncu_summary.md" % RND, "profiles/traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else ""
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
def tobytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
def tous(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1)
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("__gbs", "GB/s"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 pipe %")]
lines = []
cls_bytes = 0.0
ray_rows = [r for r in rows[2:] if "classify2_kernel" in col(r, "Kernel Name")]
ray_rows = ray_rows[:2]      # one step = 2 directions, one launch each
for r in ray_rows:
    cls_bytes += tobytes(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")])
    cls_bytes += tobytes(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")])
for r in rows[2:]:
    name = col(r, "Kernel Name").split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    vals = []
    for k, _ in keys:
        v, u = (col(r, k), units[hdr.index(k)]) if k in hdr else ("", "")
        if k == "__gbs":
            t_us = tous(col(r, "gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")])
            by = tobytes(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")]) + tobytes(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")])
            vals.append("%.0f (%.0f%%)" % (by / t_us / 1e3, 100 * by / t_us / 1e3 / 6558.1) if t_us else "")
        elif k.startswith("gpu__time"):
            vals.append("%.1f us" % tous(v, u))
        elif k.startswith("dram__bytes"):
            vals.append("%.1f MB" % (tobytes(v, u) / 1e6))
        else:
            vals.append("%.1f" % num(v) if v else "")
    lines.append("| %s | %s |" % (name[:48], " | ".join(vals)))
# the capture holds ONE serial step: 2 x (scan, hit, finish) [+ lazy second passes]
json.dump({"config": "c3", "classify_dram_bytes_per_step": int(cls_bytes),
           "how": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the two classify2_kernel launches of one step, "
                  "ncu --set full --clock-control none, scripts/stage_times.py c3 --serial (profiles/%s_ncu_summary.md)" % RND},
          open(out_json, "w"), indent=1)
# launch list
lrows = list(csv.reader(open(launch_csv)))
hi = [i for i, r in enumerate(lrows) if r and r[0] == "ID"][0]
lh = lrows[hi]; data = lrows[hi + 1:]
kn, mv, mu = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-56:]
    agg.setdefault(name, []).append(tous(r[mv], r[mu]))
tot = sum(sum(v) for v in agg.values())
with open(out_md, "w") as f:
    f.write("# Round " + RND[1:].lstrip("0") + " -- ncu evidence (B200, C3 = icosphere k=8 1,310,720 tris vs torus 1024x512 1,048,576 tris)\n\n")
    f.write("Everything here was measured UNDER ncu (cold caches, kernels serialised, streams not overlapped):\n"
            "compare SHARES, never absolutes; bench values are never taken from these runs.\n\n")
    f.write("## Launch list of `SB_GRAPHS=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (graphs off so that every kernel is listed)\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (raw list: profiles/" + RND + "_launches_bench.csv;\n"
            "C3: 3 resident steps + 9 host-buffer steps; the `other_configs` leg: 13 steps each of C2 and of C4 at its stated size (8.2 M candidate pairs: most of the\n"
            "broad_phase / predicate / classify2 time of this list); the `next_rows` leg: 4 repetitions of the widened rows (cuts_*, rank_*, uncut_*, halfedge_*, cc_*);\n"
            "fp64_peak_kernel = the FP64 issue-rate microbenchmark of the bench line; torch kernels = L2 flush / result copies)\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, len(v), sum(v), 100 * sum(v) / tot))
    f.write("| **all** | %d | %.1f | 100%% |\n\n" % (sum(len(v) for v in agg.values()), tot))
    f.write("## Per-kernel counters of one serial step (`SB_GRAPHS=0 scripts/stage_times.py c3 2 --serial`, the launches of the second step)\n")
    f.write("`ncu --section SpeedOfLight,MemoryWorkloadAnalysis,ComputeWorkloadAnalysis,Occupancy,LaunchStats,WarpStateStats,SchedulerStats,InstructionStats`\n"
            "`+ dram / L1-pipe / fp64 metrics, --clock-control none`; execution order (the window may start inside a step):\n"
            "build(A), build(B), broad phase, predicate, hit-key sort, classify A-in-B, classify B-in-A; torch kernels = the script's own result checks.\n"
            "`L1 pipe %` = l1tex__data_pipe_lsu_wavefronts (the busiest unit of the classifier).  The source-level capture of the classifier\n"
            "(`--set full --import-source on`) is summarised in profiles/" + RND + "_classify_lines.md; `GB/s` = (dram rd + dram wr) / time against the measured 6,558 GB/s copy rate.\n\n")
    f.write("| kernel | " + " | ".join(n for _, n in keys) + " |\n|---|" + "---:|" * len(keys) + "\n")
    f.write("\n".join(lines) + "\n\n")
    f.write("Classification kernel, DRAM bytes per step (two launches): **%.0f MB** vs %.0f MB algorithmic\n"
            "(SURVEY 8d: 25Q + 104 nT + 108 C with Q = 2,359,296, C = 7.21 M exact candidates under the lazy vote) -> profiles/traffic.json\n"
            % (cls_bytes / 1e6, (25 * 2359296 + 104 * 2359296 + 108 * 7.21e6) / 1e6))
print(open(out_md).read()[:3000])
