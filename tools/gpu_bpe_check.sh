#!/bin/bash
# k_bpe development check: the parity tests that exercise the merge rounds, then per-config kernel times.
# Usage: bash tools/gpu_bpe_check.sh [tag]
TAG=${1:-bpe}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "merge_rounds or golden or tiktoken or fuzz or edge or special or cfg1 or cfg4 or cfg5 or cfg3_mixed or custom_vocab or sentencepiece_golden or sentencepiece_fuzz" \
    > gpurun_out/pytest_bpe_${TAG}.log 2>&1
tail -12 gpurun_out/pytest_bpe_${TAG}.log
MB=${MB:-64} timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1
cat gpurun_out/cfgs_${TAG}.txt | tail -10
