# per-kernel times of the build with and without depth slabs (serialised under ncu: compare shares)
mkdir -p gpurun_out
for sb in 0 3; do
  SB_GRID_SLABS=$sb SB_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/slab_launch_$sb.csv \
     python scripts/stage_times.py c3 2 --serial > gpurun_out/slab_ncu_$sb.log 2>&1
  tail -2 gpurun_out/slab_ncu_$sb.log | cut -c1-300
done
