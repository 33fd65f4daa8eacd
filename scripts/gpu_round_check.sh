# Round check on the GPU box: parity tests, whole combine() through the C++ mirror, half-edge stage timing.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python scripts/combine_times.py c3 3 > gpurun_out/combine_c3.log 2>&1; tail -4 gpurun_out/combine_c3.log
timeout 300 python scripts/combine_times.py c2 3 > gpurun_out/combine_c2.log 2>&1; tail -2 gpurun_out/combine_c2.log
