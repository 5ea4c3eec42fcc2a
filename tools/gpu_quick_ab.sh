#!/bin/bash
# quick check of a change: build, some GPU tests, per-config times.  Usage: bash tools/gpu_quick_ab.sh <tag> "<pytest -k>" [cfg ...]
TAG=${1:-x}; K=${2:-golden}; shift 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python tools/gpu_cfgs.py "$@" > gpurun_out/cfgs_${TAG}.txt 2>&1; cut -c1-260 gpurun_out/cfgs_${TAG}.txt
