# why is bench.py's e2e leg slower than scripts/e2e_quick.py?  one ingredient at a time
python scripts/e2e_quick.py 30 2>&1 | tail -1
E2E_CLOSE_INSIDE=1 python scripts/e2e_quick.py 30 2>&1 | tail -1
E2E_CLOSE_INSIDE=1 E2E_EXT=1 python scripts/e2e_quick.py 30 2>&1 | tail -1
E2E_CLOSE_INSIDE=1 E2E_EXT=1 E2E_SAMPLER=1 python scripts/e2e_quick.py 30 2>&1 | tail -1
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err
python - <<PY
import json
d = json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("bench ms_per_step", round(d["ms_per_step"], 4), "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("ms_per_step_new_meshes_every_step"))
PY
