# compile-time variants (scripts/variants.py) against the product build: stage times at C3
V=solidboolean_b200/lib/variants
for v in default minb5 minb4 r256x16 default; do
  if [ $v = default ]; then L=""; else L=$PWD/$V/libsb_$v.so; fi
  echo "== $v"
  SB_LIB_PATH=$L python scripts/stage_times.py c3 8 2>&1 | tail -4 | head -2 | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "update or host_outputs or bundled or cell_borders or c3_every" 2>&1 | tail -3
