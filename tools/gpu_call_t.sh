#!/bin/bash
# slab_direct (migrants packed straight into the neighbours' buffers): slab tests, then A/B of the option at N ranks
N=${1:-2}; TAG=${2:-r02t$N}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_slabs.py -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
bash tools/gpu_call_n2.sh $N ${TAG} "--steps 40" "--steps 40 --option slab_direct=0"
