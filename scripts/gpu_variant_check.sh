# where do the resident bench step's extra microseconds come from?  stage times with the L2 flushed between steps, per variant
V=solidboolean_b200/lib/variants
for v in "default 2" "default 0" "nohist 2" "noguard 2"; do set -- $v
  if [ $1 = default ]; then L=""; else L=$PWD/$V/libsb_$1.so; fi
  echo "== $1 slabs $2"
  SB_LIB_PATH=$L SB_GRID_SLABS=$2 python scripts/stage_times.py c3 8 --flush 2>&1 | tail -4 | head -2 | cut -c1-200
  SB_LIB_PATH=$L SB_GRID_SLABS=$2 python scripts/stage_times.py c3 8 2>&1 | tail -3 | head -1 | cut -c1-200
done
