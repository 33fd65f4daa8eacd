"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py rep [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_bytes.sum','l1tex__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','launch__grid_size','launch__block_size',
 'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_xu.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
 'sm__sass_inst_executed_op_local_ld.sum','sm__sass_inst_executed_op_local_st.sum','lts__t_sectors_srcunit_tex_op_atom.sum','lts__t_sectors_srcunit_tex_op_red.sum']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if filt not in name: continue
    print('---', name[:70])
    for k in keys:
        if k in hdr:
            print('  %-88s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
