set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for e in 1 0; do
  if [ $e = 1 ]; then export SB_NO_SMALL_SORT=1; else unset SB_NO_SMALL_SORT; fi
  python scripts/stage_times.py c2 10 2>&1 | tail -3 | head -1 | cut -c1-260
  python scripts/stage_times.py c3 8 2>&1 | tail -3 | head -1 | cut -c1-260
done
