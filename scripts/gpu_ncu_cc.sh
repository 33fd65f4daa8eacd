set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cc_hook -c 2 -o gpurun_out/cc_hook -f \
    python scripts/halfedge_times.py c3 1 --no-ref > gpurun_out/ncu_cc.log 2>&1; tail -2 gpurun_out/ncu_cc.log
ls -la gpurun_out/
