#!/bin/bash
# Dev tool: one --set full capture of the build kernels of mesh A at C3 (graphs off so that every kernel is a launch)
SB_GRAPHS=0 timeout 900 ncu --set full --import-source on --clock-control none \
   -k regex:"grid_fill|leaf_gather|tri_prepare|onesweep|hist_kernel|inclusive_scan|bounds_pad|tree_build" --launch-skip 44 -c 11 -o gpurun_out/build_full -f \
   python scripts/build_times.py c3 > gpurun_out/build_full.log 2>&1
ls -la gpurun_out/build_full.ncu-rep
