#!/bin/bash
# multi-GPU call: slab tests + strong-scaling bench at N ranks
N=${1:-2}; TAG=${2:-r02n$N}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_slabs.py tests/test_gpu_brownian.py -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("N=$N ms/step %.4f value %.3e"%(d["ms_per_step"], d["value"]), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_checked"], d["parity"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("failed", e)
PY
