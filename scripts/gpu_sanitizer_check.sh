# compute-sanitizer over the parity tests of the widened rows (small cases only).
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_contexts.py tests/test_gpu_halfedge.py -m "gpu and not slow" -q -x > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_contexts.py tests/test_gpu_halfedge.py -m "gpu and not slow" -q -x -k "fixtures or edge" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_contexts.py tests/test_gpu_halfedge.py -m "gpu and not slow" -q -x -k "fixtures" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -6 gpurun_out/sanitizer_initcheck.log
