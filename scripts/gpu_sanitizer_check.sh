# compute-sanitizer over small parity tests: the widened rows (contexts, half-edge map, face groups) and the round-2 kernels
# (balanced classifier, shards, batch meshes, sb_comm).  racecheck hazards are checked against an allow-list
# (scripts/racecheck_allow.py): only the lock-free shared-memory union-find of cc_tile_kernel may appear.
set -x
mkdir -p gpurun_out
W="tests/test_gpu_contexts.py tests/test_gpu_halfedge.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $W -m "gpu and not slow" -q -x > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_shard.py tests/test_gpu_comm.py -m "gpu and not slow" -q -x \
    -k "bundled or small_and_ragged or single_triangle or cell_borders or ragged_jobs or degenerate or bad_arguments or update" > gpurun_out/sanitizer_memcheck2.log 2>&1; echo "memcheck (round 2 kernels) rc=$?"; tail -4 gpurun_out/sanitizer_memcheck2.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_halfedge.py tests/test_gpu_comm.py tests/test_gpu_parity.py -m "gpu and not slow" -q -x \
    -k "face_groups or comm_equals or hit_edge or parked or far_apart or bad_arguments" > gpurun_out/sanitizer_memcheck3.log 2>&1; echo "memcheck (flood, comm, edge tags, parked meshes) rc=$?"; tail -4 gpurun_out/sanitizer_memcheck3.log
timeout 1500 compute-sanitizer --tool racecheck python -m pytest $W -m "gpu and not slow" -q -x -k "fixtures or edge or fragments" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
python scripts/racecheck_allow.py gpurun_out/sanitizer_racecheck.log; echo "racecheck (widened rows) allow-list rc=$?"
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_shard.py -m "gpu and not slow" -q -x \
    -k "bundled or small_and_ragged or cell_borders or ragged_jobs or degenerate or many_layers" > gpurun_out/sanitizer_racecheck2.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck2.log
python scripts/racecheck_allow.py gpurun_out/sanitizer_racecheck2.log; echo "racecheck (round 2 kernels) allow-list rc=$?"
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest $W -m "gpu and not slow" -q -x -k "fixtures" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/sanitizer_initcheck.log
