#!/bin/bash
# programmatic dependent launch between the kernel nodes of a pass's graph (default) against ordinary edges (SPL_PDL=0)
TAG=${1:-x}; K=${2:-device}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}.log
for m in 1 0 1; do
  SPL_PDL=$m timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}_pdl$m.txt 2>&1; echo "SPL_PDL=$m"; grep "best" gpurun_out/cfgs_${TAG}_pdl$m.txt | cut -c1-200
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --only-headline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json | cut -c1-1500
