#!/bin/bash
# multi-GPU variants: bench at N ranks with different runtime options
N=${1:-2}; TAG=${2:-r02n$N}; shift; shift
mkdir -p gpurun_out
i=0
for opts in "$@"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $N $opts > gpurun_out/${TAG}_v$i.json 2> gpurun_out/${TAG}_v$i.err; echo "variant $i [$opts] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_v$i.json").read().strip().splitlines()[-1])
    print("  N=$N ms/step %.4f warm %.4f value %.3e"%(d["ms_per_step"], d["ms_per_step_warm_l2"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_checked"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("failed", e, open("gpurun_out/${TAG}_v$i.err").read()[-800:])
PY
done
