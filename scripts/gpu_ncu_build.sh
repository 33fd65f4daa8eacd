set -x
mkdir -p gpurun_out
SB_GRAPHS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"leaf_gather|grid_fill" -c 4 -o gpurun_out/build_kernels -f \
    python scripts/stage_times.py c3 1 > gpurun_out/ncu_build.log 2>&1; tail -2 gpurun_out/ncu_build.log
