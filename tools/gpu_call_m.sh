#!/bin/bash
# multi-GPU check after the ColSet / ABI changes: slab tests, strong-scaling bench of C2-16M and of C3 / C4 at N ranks
N=${1:-2}; TAG=${2:-r02m$N}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_slabs.py -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
run() { wl=$1; port=$2; shift; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload $wl "$@" > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("  $wl N=$N ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_checked"])
except Exception as e: print("failed", e, open("gpurun_out/${TAG}_$wl.err").read()[-800:])
PY
}
run c2_16m 29511 --steps 40
run c3 29512 --steps 10 --warmup 3
run c4 29513 --steps 10 --warmup 3
