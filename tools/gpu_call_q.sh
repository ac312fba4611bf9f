#!/bin/bash
# own-message shortcut (fgpu_launch.mirror): GPU suite, then A/B on C2 (16M, 1M), C3, C5'
mkdir -p gpurun_out
TAG=${1:-r02q}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
AB_WORKLOAD=c2_16m python tools/ab_plugins.py Brownian_16M:mirror=0 Brownian_16M 2>&1 | tee gpurun_out/${TAG}_ab16.log
python tools/ab_plugins.py Brownian:mirror=0 Brownian 2>&1 | tee gpurun_out/${TAG}_ab1m.log
python tools/ab_model.py c3 Boids_4M:mirror=0 Boids_4M 2>&1 | tee gpurun_out/${TAG}_ab_c3.log
python tools/ab_model.py c5p BirthDeath:mirror=0 BirthDeath 2>&1 | tee gpurun_out/${TAG}_ab_c5p.log
python tools/ab_model.py c4_2m SPH_2M:mirror=0 SPH_2M 2>&1 | tee gpurun_out/${TAG}_ab_c4.log
