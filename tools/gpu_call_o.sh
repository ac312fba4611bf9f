#!/bin/bash
# ncu --set full of the dominant kernel of the other BASELINE configs (after a plain run of the same command)
mkdir -p gpurun_out
TAG=${1:-r02o}
cap() { wl=$1; rx=$2; skip=$3
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
  $CMD > gpurun_out/${TAG}_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -o gpurun_out/${TAG}_$wl $CMD > gpurun_out/${TAG}_ncu_$wl.log 2>&1
  echo "$wl rc=$?"
}
cap c3 "fgpu_k_inputdata" 4
cap c4_2m "fgpu_k_computeForce" 4
cap c5_2m "fgpu_k_resolve_forces" 4
