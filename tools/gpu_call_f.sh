#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02f}
for wl in c3 c4_2m c4 c5p; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
  tail -c 400 gpurun_out/${TAG}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "ms/step %.4f value %.3e frac %.3f"%(d["ms_per_step"], d["value"], d["roofline"]["frac"]), d["roofline"]["stage_ms"], "parity", d["parity_checked"], "pop", d["population"], "cpu", d.get("cpu_baseline",{}).get("value"), "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("$wl failed", e)
PY
done
python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench16.json 2> gpurun_out/${TAG}_bench16.err; echo "bench16 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench16.json").read().strip().splitlines()[-1])
print("16M ms/step", d["ms_per_step"], "stages", d["roofline"]["stage_ms"], "parity", d["parity_checked"], "e2e", d["e2e"]["value"], "serial", d["e2e"].get("serial",{}).get("value"))
PY
