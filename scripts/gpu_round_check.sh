# Round check on the GPU box: smoke, parity tests, bench lines (ours + reference arm), launch list, widened rows.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; wc -l gpurun_out/bench_c3.json; tail -c 600 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
SB_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 300 gpurun_out/ncu_bench.log
timeout 300 python scripts/combine_times.py c3 3 > gpurun_out/combine_c3.log 2>&1; tail -2 gpurun_out/combine_c3.log | cut -c1-400
