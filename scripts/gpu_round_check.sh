# Round check on the GPU box: parity tests, half-edge stage timing (+ launch list), bench line.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python scripts/halfedge_times.py c3 5 > gpurun_out/halfedge_c3.log 2>&1; tail -3 gpurun_out/halfedge_c3.log
timeout 200 python scripts/halfedge_times.py c2 5 --no-ref > gpurun_out/halfedge_c2.log 2>&1; tail -1 gpurun_out/halfedge_c2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_halfedge.csv \
    python scripts/halfedge_times.py c3 2 --no-ref > gpurun_out/ncu_halfedge.log 2>&1; tail -2 gpurun_out/ncu_halfedge.log
