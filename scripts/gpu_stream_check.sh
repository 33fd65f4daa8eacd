# host-output front end + optimistic enqueue + streamed classification: parity tests, stage times, e2e step, bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_outputs or streamed or front_end or update" 2>&1 | tail -8 > gpurun_out/stream_pytest.log; cat gpurun_out/stream_pytest.log
python scripts/stage_times.py c3 6 2>&1 | tail -3 | head -1 | cut -c1-330
SB_OPTIMISTIC=0 python scripts/e2e_quick.py 30 2>&1 | tail -1
E2E_SPANS=1 python scripts/e2e_quick.py 30 2>&1 | tail -12
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/stream_bench_1.json 2> gpurun_out/stream_bench_1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/stream_bench_1.json").read().strip().splitlines()[-1])
print("bench ms_per_step", round(d["ms_per_step"], 4), "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("ms_per_step_new_meshes_every_step"), d["stage_ms"], d.get("parity_vs_cpu"), d["roofline"]["frac"])
PY
tail -3 gpurun_out/stream_bench_1.err
