#!/bin/bash
# quick GPU check: parity tests + bench (no ncu).  Usage: bash tools/gpu_quick.sh [tag]
TAG=${1:-q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -15 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
