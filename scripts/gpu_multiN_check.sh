# N GPUs: the driver's launch lines for C3 and C5
set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; wc -l gpurun_out/bench_c3_n$N.json; tail -3 gpurun_out/bench_c3_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; wc -l gpurun_out/bench_c5_n$N.json; tail -3 gpurun_out/bench_c5_n$N.err
python - <<PY
import json
for f in ("bench_c3_n$N", "bench_c5_n$N"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), (d.get("multi_gpu") or {}).get("parity_vs_single_gpu"), d.get("parity"))
PY
