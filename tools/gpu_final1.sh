#!/bin/bash
# final single-GPU pass: whole GPU suite, the bench line (+ reference arm), ingestion timings.  Usage: bash tools/gpu_final1.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -5 gpurun_out/pytest_gpu_${TAG}.log
bash tools/gpu_multi.sh ${TAG} 1 10 3
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; tail -c 600 gpurun_out/bench_ref_${TAG}.json
timeout 600 python tools/gpu_parquet.py > gpurun_out/parquet_${TAG}.txt 2>&1; tail -20 gpurun_out/parquet_${TAG}.txt
