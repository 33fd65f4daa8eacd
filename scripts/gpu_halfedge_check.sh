# Half-edge stage on the GPU box: its parity tests, live timing, launch list, per-kernel counters.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halfedge.py -m gpu -q 2>&1 | tail -8
timeout 300 python scripts/halfedge_times.py c3 5 > gpurun_out/halfedge_c3.log 2>&1; tail -2 gpurun_out/halfedge_c3.log
timeout 200 python scripts/halfedge_times.py c2 5 --no-ref > gpurun_out/halfedge_c2.log 2>&1; tail -1 gpurun_out/halfedge_c2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_halfedge.csv \
    python scripts/halfedge_times.py c3 2 --no-ref > gpurun_out/ncu_halfedge.log 2>&1; tail -1 gpurun_out/ncu_halfedge.log
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_requests_srcunit_tex_op_atom_dot_cas.sum \
    --clock-control none -k regex:"uncut|halfedge|cc_|onesweep|hist_kernel" --launch-skip 20 -c 60 -o gpurun_out/halfedge_kernels -f \
    python scripts/halfedge_times.py c3 1 --no-ref > gpurun_out/ncu_halfedge_kernels.log 2>&1; tail -1 gpurun_out/ncu_halfedge_kernels.log
