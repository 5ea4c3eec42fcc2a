#!/bin/bash
# Parquet / Arrow ingestion: GPU tests + timings (snappy out of shared memory against the plain decoder).  Usage: bash tools/gpu_n4.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 900 python -m pytest tests -m gpu -x -q -k "parquet or arrow or jsonl" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -15 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python tools/gpu_parquet.py > gpurun_out/parquet_${TAG}.txt 2>&1; cat gpurun_out/parquet_${TAG}.txt | tail -20
SPL_PQ_SNAPPY_RING=0 timeout 600 python tools/gpu_parquet.py snappy > gpurun_out/parquet_${TAG}_plain.txt 2>&1; cat gpurun_out/parquet_${TAG}_plain.txt | tail -20
