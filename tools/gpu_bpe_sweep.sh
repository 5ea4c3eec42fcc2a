#!/bin/bash
# k_bpe: parity at the default window class, then kernel times of cfg4 / cfg5 for several window classes.
TAG=${1:-sw}
bash tools/gpu_bpe_check.sh $TAG
for W in ${WINS:-2 3}; do
  echo "== SPL_BPE_WIN_CLS=$W"
  SPL_BPE_WIN_CLS=$W MB=64 timeout 300 python tools/gpu_cfgs.py cfg4 cfg5 2>&1 | grep -v "^ *$" | sed 's/k_mark_docs.*k_bpe_long/k_bpe_long/'
done | tee gpurun_out/sweep_${TAG}.txt
