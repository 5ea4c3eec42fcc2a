#!/bin/bash
# ncu --set full of k_bpe on cfg4 / cfg5 (tools/gpu_cfgs.py).  Usage: bash tools/gpu_ncu_bpe.sh <tag> [cfg ...]
TAG=${1:-x}; shift
for c in ${@:-cfg4 cfg5}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bpe' -s 4 -c 2 -f -o gpurun_out/prof_bpe_${c}_${TAG} \
      python tools/gpu_cfgs.py $c > gpurun_out/ncu_bpe_${c}_${TAG}.log 2>&1
  tail -n 2 gpurun_out/ncu_bpe_${c}_${TAG}.log
done
