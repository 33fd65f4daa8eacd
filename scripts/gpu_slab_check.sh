# Depth slabs in the ray grids: parity tests of everything that reads the grids, then stage times per slab setting.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_shard.py tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/slab_pytest.log; cat gpurun_out/slab_pytest.log
for sb in 0 1 2 3; do
  echo "== SB_GRID_SLABS=$sb"
  SB_GRID_SLABS=$sb timeout 200 python scripts/stage_times.py c3 6 2>&1 | tail -4 | cut -c1-700
  SB_GRID_SLABS=$sb timeout 200 python scripts/stage_times.py c3 4 --serial 2>&1 | tail -3 | head -2 | cut -c1-500
done > gpurun_out/slab_times.log 2>&1
cat gpurun_out/slab_times.log
for sb in 0 3; do
  echo "== c4k8 SB_GRID_SLABS=$sb"
  SB_GRID_SLABS=$sb timeout 200 python scripts/stage_times.py c4k8 4 2>&1 | tail -3 | head -1 | cut -c1-500
done > gpurun_out/slab_times_c4.log 2>&1
cat gpurun_out/slab_times_c4.log
