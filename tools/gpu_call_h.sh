#!/bin/bash
# GPU call: tests, A/B of the if-conversion budget on C2 (1M and 16M), default bench, full ncu of count_neighbours at 16M
mkdir -p gpurun_out
TAG=${1:-r02h}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
AB_WORKLOAD=c2_16m python tools/ab_plugins.py Brownian_16M_ifc0 Brownian_16M 2>&1 | tee gpurun_out/${TAG}_ab16.log
python tools/ab_plugins.py Brownian_ifc0 Brownian 2>&1 | tee gpurun_out/${TAG}_ab1m.log
python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench16.json 2> gpurun_out/${TAG}_bench16.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench16.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity"
ncu --set full --clock-control none --import-source on -k regex:"count_neighbours" -s 4 -c 1 -o gpurun_out/${TAG}_cn16 $CMD > gpurun_out/${TAG}_ncu_cn.log 2>&1
echo "ncu cn rc=$?"
