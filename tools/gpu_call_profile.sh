#!/bin/bash
# end-of-round profile at C2-16M: launch list + full ncu of the top kernels (after a plain run of the same command)
mkdir -p gpurun_out
TAG=${1:-r02p}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity"
$CMD > gpurun_out/${TAG}_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches16.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"count_neighbours" -s 4 -c 1 -o gpurun_out/${TAG}_cn16 $CMD > gpurun_out/${TAG}_ncu_cn.log 2>&1
echo "ncu cn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_tile_hist|k_radix_scatter|k_gather_pbm|k_row_scan" -s 40 -c 5 -o gpurun_out/${TAG}_build16 $CMD > gpurun_out/${TAG}_ncu_build.log 2>&1
echo "ncu build rc=$?"
