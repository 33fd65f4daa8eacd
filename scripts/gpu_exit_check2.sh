# exit status of the driver's torchrun lines on two GPUs (ours, the reference arm, C5)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > /tmp/o.json 2>/tmp/o.err; echo "ours N=2 rc=$? lines=$(wc -l < /tmp/o.json)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > /tmp/r.json 2>/tmp/r.err; echo "reference N=2 rc=$? lines=$(wc -l < /tmp/r.json)"; head -c 200 /tmp/r.json; echo
python -m pytest tests/test_gpu_multi.py -m gpu -q > /tmp/t.log 2>&1; echo "test_gpu_multi rc=$?"; tail -1 /tmp/t.log
