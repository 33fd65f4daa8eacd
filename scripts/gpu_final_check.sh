# End-of-round check on the GPU box: sanitizers over the kernels touched last (grid build with depth slabs, fused histogram,
# scan look-back, guarded classifier, host-output front end), the whole GPU test suite, smoke, the bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m "gpu and not slow" -q -x \
    -k "bundled or small_and_ragged or single_triangle or cell_borders or ragged_jobs or degenerate or update or host_outputs" > gpurun_out/sanitizer_memcheck4.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck4.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m "gpu and not slow" -q -x \
    -k "bundled or small_and_ragged or cell_borders or ragged_jobs or host_outputs" > gpurun_out/sanitizer_racecheck4.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck4.log
python scripts/racecheck_allow.py gpurun_out/sanitizer_racecheck4.log; echo "racecheck allow-list rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json
python scripts/stage_times.py c3 8 2>&1 | tail -3 | head -1 | cut -c1-300
