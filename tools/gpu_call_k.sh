#!/bin/bash
# end-of-round refresh: every BASELINE config on one GPU (default bench incl. CPU baseline and reference CUDA, C2-1M, C3, C4, C5'),
# then the ncu launch list of the default bench command
mkdir -p gpurun_out
TAG=${1:-r02k}
python bench.py > gpurun_out/${TAG}_bench16.json 2> gpurun_out/${TAG}_bench16.err; echo "bench16 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref16.json 2> gpurun_out/${TAG}_ref16.err; echo "ref16 rc=$?"
python bench.py --workload c2 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err; echo "c2 rc=$?"
for wl in c3 c4_2m c4 c5p; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
done
python - <<PY
import json
for wl in ("bench16", "ref16", "c2", "c3", "c4_2m", "c4", "c5p"):
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
