# Contexts stage on the GPU box: parity tests + live timing at three sizes.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contexts.py -m gpu -q 2>&1 | tail -8
for c in c3 c4 c4k8; do timeout 300 python scripts/contexts_times.py $c 5 > gpurun_out/contexts_$c.log 2>&1; tail -2 gpurun_out/contexts_$c.log | cut -c1-400; done
