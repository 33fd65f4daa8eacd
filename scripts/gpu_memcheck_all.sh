# compute-sanitizer memcheck over every GPU test that is not marked slow (the large configurations)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m "gpu and not slow" -q -x > gpurun_out/sanitizer_memcheck_all.log 2>&1; echo "memcheck (all GPU tests that are not slow) rc=$?"; tail -6 gpurun_out/sanitizer_memcheck_all.log
