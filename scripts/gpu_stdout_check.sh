set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; wc -l gpurun_out/b1.json; head -c 120 gpurun_out/b1.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; wc -l gpurun_out/bench_c3_n2.json; head -c 120 gpurun_out/bench_c3_n2.json; echo; grep -c "NCCL version" gpurun_out/bench_c3_n2.err
