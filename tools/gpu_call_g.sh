#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02g}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
run() { n=$1; shift
  python bench.py --no-cpu-baseline --no-reference-cuda --no-parity --steps 40 "$@" > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$n.json").read().strip().splitlines()[-1])
    print("$n", "ms/step %.4f warm %.4f"%(d["ms_per_step"], d["ms_per_step_warm_l2"]), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("$n failed", e, open("gpurun_out/${TAG}_$n.err").read()[-500:])
PY
}
run pdl1_16
run pdl0_16 --option pdl=0
run pdl1_1m --workload c2
run pdl0_1m --workload c2 --option pdl=0
