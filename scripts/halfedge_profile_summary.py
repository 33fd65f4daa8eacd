"""Write profiles/r01_halfedge_summary.md from the ncu artefacts of scripts/gpu_round_check.sh:
python scripts/halfedge_profile_summary.py gpurun_out/launches_halfedge.csv gpurun_out/halfedge_kernels.ncu-rep gpurun_out/halfedge_c3.log"""
import csv, json, subprocess, sys
launch_csv, rep, log = sys.argv[1:4]
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
def scale(k, v):
    u = units[hdr.index(k)] if k in hdr else ""
    return num(v) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1)
keys = [("gpu__time_duration.sum", "time us", 1), ("dram__bytes_read.sum", "dram rd MB", 1e-6), ("dram__bytes_write.sum", "dram wr MB", 1e-6),
        ("launch__registers_per_thread", "regs", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1),
        ("lts__t_requests_srcunit_tex_op_atom_dot_cas.sum", "L2 CAS", 1)]
out = ["# Round 1 -- half-edge stage (SURVEY 8f rows 2-3) under ncu, C3 (1,305,236 + 1,044,454 uncut triangles)", "",
       "Measured UNDER ncu (cold caches, serialised): compare shares, not absolutes.  Live timings (CUDA events of the library's",
       "stage timer, `scripts/halfedge_times.py c3`, last line = best repetition + the reference's own functions on one host core):", "", "```"]
out += [l.rstrip() for l in open(log).read().strip().splitlines()[-2:]]
out += ["```", "", "## Launch list of `python scripts/halfedge_times.py c3 2 --no-ref` (`ncu --metrics gpu__time_duration.sum --clock-control none`)",
        "(builds + intersection of the set-up included; 2 repetitions x 2 meshes of the stage itself)", ""]
out += subprocess.run([sys.executable, "scripts/launch_summary.py", launch_csv], capture_output=True, text=True).stdout.strip().splitlines()
out += ["", "## Per-kernel counters of one repetition (mesh A = icosphere, then mesh B = torus; components of A, of B)",
        "`ncu --section SpeedOfLight,MemoryWorkloadAnalysis,Occupancy,WarpStateStats,LaunchStats + dram / CAS metrics, --clock-control none`", "",
        "| kernel | " + " | ".join(k[1] for k in keys) + " |", "|---|" + "---:|" * len(keys)]
tot_r = tot_w = tot_t = 0.0
for r in rows[2:]:
    name = col(r, "Kernel Name").split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:48]
    vals = []
    for k, _, f in keys:
        v = scale(k, col(r, k)) * f
        vals.append("%.1f" % v if v < 1e5 else "%d" % v)
    if "cc_" not in name:
        tot_t += scale(keys[0][0], col(r, keys[0][0])); tot_r += scale(keys[1][0], col(r, keys[1][0])); tot_w += scale(keys[2][0], col(r, keys[2][0]))
    out.append("| %s | %s |" % (name, " | ".join(vals)))
out += ["", "Half-edge map kernels (without the components): %.1f us, DRAM %.0f MB read + %.0f MB written = %.0f MB against 1,271 MB algorithmic"
        % (tot_t, tot_r / 1e6, tot_w / 1e6, (tot_r + tot_w) / 1e6),
        "(DESIGN section 4: the sort's ping-pong buffers of mesh B partly stay in the 126 MB L2).",
        "Components: the union-find runs on positions of the mesh's Morton order (cc_rank / cc_order translate), because neither",
        "generator emits faces in a spatially coherent order; in face order cc_hook alone took 340 + 115 us (0.7 + 0.9 M L2 CAS)."]
open("profiles/r01_halfedge_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[-40:]))
