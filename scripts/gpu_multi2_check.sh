# two GPUs: the multi-GPU tests, the driver's launch lines for C3 and C5, the reference arm under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_comm.py tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; wc -l gpurun_out/bench_c3_n2.json; tail -3 gpurun_out/bench_c3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --config c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; wc -l gpurun_out/bench_c5_n2.json; tail -3 gpurun_out/bench_c5_n2.err
python - <<PY
import json
for f in ("bench_c3_n2", "bench_c5_n2"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d.get("stage_ms"), (d.get("multi_gpu") or {}).get("parity_vs_single_gpu"), d.get("parity"))
PY
