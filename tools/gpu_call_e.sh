#!/bin/bash
mkdir -p gpurun_out
TAG=r02e
run() { # name, args...
  n=$1; shift
  python bench.py --no-cpu-baseline --no-reference-cuda --no-parity "$@" > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$n.json").read().strip().splitlines()[-1])
    print("$n", "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print("$n failed", e, open("gpurun_out/${TAG}_$n.err").read()[-500:])
PY
}
run base16
run l2f32_16 --option l2_fetch=32
run l2f128_16 --option l2_fetch=128
run rb8_16 --option radix_bits=8
run rb11_16 --option radix_bits=11
run base1m --workload c2
run l2f32_1m --workload c2 --option l2_fetch=32
