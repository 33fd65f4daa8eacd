#!/bin/bash
# Round-2 evidence run on the GPU box (everything lands in gpurun_out/, which must stay below 64 MiB):
#   bench : bench lines -- ours at C3 (with cpu_baseline, other_configs, next_rows), the reference arm, C5
#   ncu   : ncu launch list of the bench command, per-kernel counters of one serial step, source-level capture of the classifier
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
if [ "$1" = "bench" ]; then
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 400 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json
fi
if [ "$1" = "ncu" ]; then
SB_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 200 gpurun_out/ncu_bench.log
SB_GRAPHS=0 timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section Occupancy \
    --section LaunchStats --section WarpStateStats --section SchedulerStats --section InstructionStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sectors_srcunit_tex_op_atom.sum,lts__t_sectors_srcunit_tex_op_red.sum \
    --clock-control none --launch-skip 37 -c 37 -o gpurun_out/step_counters -f \
    python scripts/stage_times.py c3 2 --serial > gpurun_out/step_counters.log 2>&1; ls -la gpurun_out/step_counters.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:classify2 -c 2 -o gpurun_out/cls_lines -f \
    python scripts/stage_times.py c3 1 --serial > gpurun_out/cls_lines.log 2>&1; ls -la gpurun_out/cls_lines.ncu-rep
du -sh gpurun_out
fi
