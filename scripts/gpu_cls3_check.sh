#!/bin/bash
# Dev tool: parity tests of the classifier + variant timings (serial and overlapped) at C3 + one source-level ncu capture
python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -4
bash scripts/cmp_variants.sh c3 --serial
bash scripts/cmp_variants.sh c3 --overlap
bash scripts/cmp_variants.sh c2 --overlap
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:classify2 -c 2 -o gpurun_out/cls3 -f \
    python scripts/stage_times.py c3 1 --serial > gpurun_out/cls3_ncu.log 2>&1
ls -la gpurun_out/cls3.ncu-rep
fi
