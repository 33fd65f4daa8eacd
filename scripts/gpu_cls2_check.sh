#!/bin/bash
# Dev tool: parity tests, variant timings and one source-level ncu capture of the balanced classifier
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
bash scripts/cmp_variants.sh c3 --serial
bash scripts/cmp_variants.sh c3 --overlap | head -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:classify2 -c 2 -o gpurun_out/cls2 -f \
    python scripts/stage_times.py c3 1 --serial > gpurun_out/cls2_ncu.log 2>&1
ls -la gpurun_out/cls2.ncu-rep
