#!/bin/bash
# Round-2 evidence: GPU tests, per-config kernel times, ncu --set full of the merge kernels on cfg4 / cfg5.
# Usage: bash tools/gpu_r2a.sh <tag> [pytest -k expression]
TAG=${1:-r02a}
K=${2:-}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu_${TAG}.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
fi
tail -15 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1
tail -12 gpurun_out/cfgs_${TAG}.txt
echo "--- SPL_PROBE_BULK=0 (cfg2 cfg5)"; SPL_PROBE_BULK=0 timeout 600 python tools/gpu_cfgs.py cfg2 cfg5 > gpurun_out/cfgs_${TAG}_nobulk.txt 2>&1; grep -A1 "^cfg" gpurun_out/cfgs_${TAG}_nobulk.txt
echo "--- SPL_NO_DEDUP=1 (cfg4)"; SPL_NO_DEDUP=1 timeout 600 python tools/gpu_cfgs.py cfg4 > gpurun_out/cfgs_${TAG}_nodedup.txt 2>&1; grep -A1 "^cfg" gpurun_out/cfgs_${TAG}_nodedup.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
bash tools/gpu_ncu_bpe.sh ${TAG}
ls -la gpurun_out | tail -8
