# N GPUs: the driver's launch line for C3 only
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_c3_n$N.json)"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c3_n$N.json").read().strip().splitlines()[-1])
print("bench_c3_n$N ms_per_step", round(d["ms_per_step"], 4), "instrumented", round(d["ms_per_step_instrumented"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), (d.get("multi_gpu") or {}).get("parity_vs_single_gpu", {}).get("flags_identical"))
PY
