for f in solidboolean_b200/lib/libsolidboolean_b200.so solidboolean_b200/lib/variants/libsb_t*.so; do
  echo "== $f"
  SB_LIB_PATH=$PWD/$f timeout 200 python scripts/halfedge_times.py c3 5 --no-ref 2>&1 | tail -1 | cut -c1-200
  SB_LIB_PATH=$PWD/$f timeout 200 python scripts/halfedge_times.py c2 5 --no-ref 2>&1 | tail -1 | cut -c1-200
done
