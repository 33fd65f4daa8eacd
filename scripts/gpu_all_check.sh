#!/bin/bash
# Dev tool: the whole GPU test suite + stage timings at C3 / C2 / C4k8
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in c3 c2 c4k8; do python scripts/stage_times.py $c 5 2>&1 | grep '"it": 4' | cut -c1-330; done
python scripts/stage_times.py c3 5 --serial 2>&1 | grep '"it": 4' | cut -c1-330
