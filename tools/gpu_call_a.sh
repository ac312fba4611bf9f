#!/bin/bash
# GPU call: tests, the default bench (C2-16M), launch list and full ncu of the top kernels at 16M
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gputests.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r02a_bench16.json 2> gpurun_out/r02a_bench16.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench16.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity"
$CMD > gpurun_out/r02a_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches16.csv $CMD > gpurun_out/r02a_ncu_l.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"count_neighbours|k_gather_pbm|k_radix_scatter|k_tile_hist" -s 24 -c 6 -o gpurun_out/r02a_prof16 $CMD > gpurun_out/r02a_ncu_full.log 2>&1
echo "ncu full rc=$?"
