#!/bin/bash
# k_bpe_long beside k_bpe in the pass's graph (default) against the plain chain (SPL_BPE_OVERLAP=0).  Usage: bash tools/gpu_overlap_ab.sh <tag> "<pytest -k>"
TAG=${1:-x}; K=${2:-device}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1; grep "best" gpurun_out/cfgs_${TAG}.txt
SPL_BPE_OVERLAP=0 timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}_chain.txt 2>&1; grep "best" gpurun_out/cfgs_${TAG}_chain.txt
timeout 600 python tools/gpu_parquet.py "snappy" > gpurun_out/parquet_${TAG}.txt 2>&1; tail -16 gpurun_out/parquet_${TAG}.txt
