#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -5 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_probe|k_emit|k_pretok_fast|k_bpe' -s 12 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
