#!/bin/bash
# ncu --set full capture of some kernels of the bench step.  Usage: bash tools/gpu_ncu.sh <kernel-regex> <tag> <skip> <count> [extra bench args]
K=${1:-k_probe}; TAG=${2:-x}; SKIP=${3:-4}; CNT=${4:-1}; shift 4
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${TAG}.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -n 3 gpurun_out/ncu_full_${TAG}.log
