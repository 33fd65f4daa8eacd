for v in "2 1.0" "0 1.0" "3 1.0" "2 0.7" "2 1.4"; do set -- $v
  SB_GRID_SLABS=$1 SB_GRID_BETA=$2 timeout 300 python bench.py --config c5 --steps 6 --warmup 2 --no-cpu-baseline > /tmp/c5.json 2> /tmp/c5.err
  python - <<PY
import json
d = json.loads(open("/tmp/c5.json").read().strip().splitlines()[-1])
print("slabs $1 beta $2: ms_per_step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["stage_ms"], d.get("parity", {}).get("identical"))
PY
done
