#!/bin/bash
# quick round: build, a few GPU tests, per-config times, one ncu capture.  Usage: bash tools/gpu_r2b.sh <tag> "<pytest -k>" <cfg> <kernel regex> <skip> <count>
TAG=${1:-x}; K=${2:-golden}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1; cat gpurun_out/cfgs_${TAG}.txt | cut -c1-300
if [ -n "$3" ]; then bash tools/gpu_ncu_one.sh $TAG $3 "$4" ${5:-2} ${6:-1}; fi
