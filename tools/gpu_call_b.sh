#!/bin/bash
# GPU call: tests, short bench at 16M and 1M, full ncu of count_neighbours at 16M
mkdir -p gpurun_out
TAG=${1:-r02b}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench16.json 2> gpurun_out/${TAG}_bench16.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench16.json").read().strip().splitlines()[-1])
print("16M ms/step", d["ms_per_step"], "stages", d["roofline"]["stage_ms"], "parity", d["parity_checked"], "e2e", d["e2e"]["value"])
PY
python bench.py --workload c2 --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench1m.json 2> gpurun_out/${TAG}_bench1m.err; echo "bench1m rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench1m.json").read().strip().splitlines()[-1])
print("1M ms/step", d["ms_per_step"], "stages", d["roofline"]["stage_ms"], "parity", d["parity_checked"])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity"
$CMD > gpurun_out/${TAG}_plain16.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"count_neighbours" -s 4 -c 1 -o gpurun_out/${TAG}_cn16 $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
