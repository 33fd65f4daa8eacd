# The driver's scaling lines on an 8-GPU box: C3 at N = 8, 4, 2 and C5 at N = 8 (one JSON line each in gpurun_out/)
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
    bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; head -c 200 gpurun_out/bench_c3_n$N.json; tail -2 gpurun_out/bench_c3_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --config c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; head -c 200 gpurun_out/bench_c5_n8.json; tail -2 gpurun_out/bench_c5_n8.err
timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -2
