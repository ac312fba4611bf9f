#!/bin/bash
# shipped-size configs C1 / C5 (bench lines) and the C2-1M line with its parity check
mkdir -p gpurun_out
TAG=${1:-r02l}
for wl in c1 c5; do
  python bench.py --workload $wl --steps 100 --warmup 10 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"; tail -c 300 gpurun_out/${TAG}_$wl.err
done
python bench.py --workload c2 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err; echo "c2 rc=$?"
python - <<PY
import json
for wl in ("c1", "c5", "c2"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % wl).read().strip().splitlines()[-1])
        print(wl, "ms/step %.4f value %.4e" % (d["ms_per_step"], d["value"]), "e2e %.3e" % d.get("e2e", {}).get("value", 0), "cpu", d.get("cpu_baseline", {}).get("value"),
              "frac", d.get("roofline", {}).get("frac"), "parity", d.get("parity_checked"), d.get("roofline", {}).get("stage_ms"), d.get("population"))
    except Exception as e:
        print(wl, "failed", e)
PY
