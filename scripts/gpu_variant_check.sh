# compile-time variants (scripts/variants.py) against the product build: stage times at C3 / C4 at its stated size
V=solidboolean_b200/lib/variants
for v in default pf1 pf2 default; do
  if [ $v = default ]; then L=""; else L=$PWD/$V/libsb_$v.so; fi
  echo "== $v"
  SB_LIB_PATH=$L python scripts/stage_times.py c3 8 2>&1 | tail -4 | head -2 | cut -c1-200
  SB_LIB_PATH=$L python scripts/stage_times.py c4k8 5 2>&1 | tail -3 | head -1 | cut -c1-200
done
