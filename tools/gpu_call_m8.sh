#!/bin/bash
# 8-GPU strong-scaling lines (no tests: tools/gpu_call_m.sh runs them at 2 ranks)
N=${1:-8}; TAG=${2:-r02m8}
mkdir -p gpurun_out
run() { wl=$1; port=$2; shift; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload $wl "$@" > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("  $wl N=$N ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_checked"], "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e: print("failed", e, open("gpurun_out/${TAG}_$wl.err").read()[-800:])
PY
}
run c2_16m 29511 --steps 40
run c4 29513 --steps 10 --warmup 3
run c3 29512 --steps 10 --warmup 3
