#!/bin/bash
# final single-GPU refresh after the own-message shortcut: default bench (both legs), C2-1M, C3, C5', C5-2M, ncu of count_neighbours + launch list
mkdir -p gpurun_out
TAG=${1:-r02s}
python bench.py > gpurun_out/${TAG}_bench16.json 2> gpurun_out/${TAG}_bench16.err; echo "bench16 rc=$?"
python bench.py --workload c2 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err; echo "c2 rc=$?"
for wl in c3 c5p c5_2m c1; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
done
python - <<PY
import json
for wl in ("bench16", "c2", "c3", "c5p", "c5_2m", "c1"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % wl).read().strip().splitlines()[-1])
        print(wl, "ms/step %.4f value %.4e" % (d["ms_per_step"], d["value"]), "e2e %.3e" % d.get("e2e", {}).get("value", 0),
              "frac", d.get("roofline", {}).get("frac"), "parity", d.get("parity_checked"), d.get("roofline", {}).get("stage_ms"))
    except Exception as e:
        print(wl, "failed", e)
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches16.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"count_neighbours" -s 4 -c 1 -o gpurun_out/${TAG}_cn16 $CMD > gpurun_out/${TAG}_ncu_cn.log 2>&1
echo "ncu cn rc=$?"
