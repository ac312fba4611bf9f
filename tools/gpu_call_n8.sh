#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02n8b}
python -m pytest tests/test_gpu_slabs.py -m gpu -x -q -k "more_slabs and 8" > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
bash tools/gpu_call_n2.sh 8 ${TAG} "--steps 40"
bash tools/gpu_call_n3.sh 8 ${TAG}m c4 c3
