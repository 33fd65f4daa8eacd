#!/bin/bash
# Dev tool: time scripts/stage_times.py with every library variant under solidboolean_b200/lib/variants
cfg=${1:-c3}; mode=${2:---serial}
for f in solidboolean_b200/lib/libsolidboolean_b200.so solidboolean_b200/lib/variants/libsb_*.so; do
  SB_LIB_PATH=$PWD/$f python scripts/stage_times.py $cfg 5 $mode 2>&1 | grep '"it": 4' | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('%-28s wall %.3f  %s' % ('$f'.split('/')[-1], d['wall_ms'], d['stages']))"
done
