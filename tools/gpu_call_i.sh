#!/bin/bash
# GPU call: A/B of the message-loop variants of tools/build_ab_variants.py on C2 (16M and 1M)
mkdir -p gpurun_out
TAG=${1:-r02i}
V="${2:-pair pipe pipe3 pair_pf1 pipe_pf1 pipe_pf2}"
AB_WORKLOAD=c2_16m python tools/ab_plugins.py $(for v in $V; do echo Brownian_16M_$v; done) 2>&1 | tee gpurun_out/${TAG}_ab16.log
python tools/ab_plugins.py $(for v in $V; do echo Brownian_$v; done) 2>&1 | tee gpurun_out/${TAG}_ab1m.log
