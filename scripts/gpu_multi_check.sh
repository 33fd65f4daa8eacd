# Multi-GPU sanity: the driver's launch line for N = $1 (default 2)
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; wc -l gpurun_out/bench_c3_n$N.json; head -c 300 gpurun_out/bench_c3_n$N.json; tail -3 gpurun_out/bench_c3_n$N.err
