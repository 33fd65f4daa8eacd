# after the bench restructure (value without stage events) and the early-preparation switch: tests with exit status, bench lines
python -m pytest tests -x -q -m gpu > /tmp/p.log 2>&1; echo "pytest tests -x -q -m gpu rc=$?"; tail -3 /tmp/p.log
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c3.json").read().strip().splitlines()[-1])
print("c3 ms_per_step", round(d["ms_per_step"], 4), "instrumented", round(d["ms_per_step_instrumented"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["e2e"].get("ms_per_step_new_meshes_every_step"), d["stage_ms"], d["roofline"]["frac"], d["gpu_launches_per_step"], d["parity_vs_cpu"]["flags_a_identical"], {k: v["ms_per_step"] for k, v in d["other_configs"].items()})
d = json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print("c5 ms_per_step", round(d["ms_per_step"], 4), "instrumented", round(d["ms_per_step_instrumented"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["parity"]["identical"])
PY
