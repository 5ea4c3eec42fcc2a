#!/bin/bash
# k_bpe_long beside k_bpe in the pass's graph: SPL_BPE_OVERLAP = 0 (chain), 1 (k_bpe_long's node first), 2 (k_bpe first on 4 blocks per SM).
TAG=${1:-x}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
for m in 2 1 0 2; do
  SPL_BPE_OVERLAP=$m timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}_ov$m.txt 2>&1; echo "overlap $m"; grep "best" gpurun_out/cfgs_${TAG}_ov$m.txt | cut -c1-200
done
bash tools/gpu_sanitize.sh 2>&1 | tail -14
