# Dev tool: product build vs the variants under solidboolean_b200/lib/variants (stage times + a parity subset).
set -x
for f in solidboolean_b200/lib/libsolidboolean_b200.so solidboolean_b200/lib/variants/libsb_*.so; do
  echo "== $f"
  SB_LIB_PATH=$PWD/$f timeout 300 python scripts/stage_times.py c3 6 2>&1 | grep '"it": [45]' | cut -c1-230
  SB_LIB_PATH=$PWD/$f timeout 300 python scripts/stage_times.py c2 6 2>&1 | grep '"it": 5' | cut -c1-230
  SB_LIB_PATH=$PWD/$f timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bundled or synthetic or c2 or c3" 2>&1 | tail -1
done
