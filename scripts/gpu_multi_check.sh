# 2-GPU sanity: the driver's launch line for N=2 (ours and the reference arm)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 1200 gpurun_out/bench_c3_n2.json; tail -5 gpurun_out/bench_c3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -c 300 gpurun_out/bench_ref_n2.json
