#!/bin/bash
# multi-GPU: model workloads (C3 / C4) in slab mode
N=${1:-2}; TAG=${2:-r02m$N}; shift; shift
mkdir -p gpurun_out
i=0
for wl in "$@"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("  $wl N=$N ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_checked"], d["parity"])
except Exception as e: print("failed", e, open("gpurun_out/${TAG}_$wl.err").read()[-1500:])
PY
done
