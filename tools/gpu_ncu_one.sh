#!/bin/bash
# ncu --set full of some kernels of one config (tools/gpu_cfgs.py).  Usage: bash tools/gpu_ncu_one.sh <tag> <cfg> <kernel regex> <skip> <count>
TAG=$1; CFG=$2; K=$3; SKIP=${4:-2}; CNT=${5:-1}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_${CFG}_${TAG} \
    python tools/gpu_cfgs.py $CFG > gpurun_out/ncu_${CFG}_${TAG}.log 2>&1
tail -n 2 gpurun_out/ncu_${CFG}_${TAG}.log
