#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/dbg_build.log 2>&1
for B in 0 1; do
  for D in 0 1; do
    echo "=== SPL_PROBE_BULK=$B SPL_NO_DEDUP=$D"; SPL_NO_DEDUP=$D SPL_PROBE_BULK=$B SPL_SYNC_EACH=1 timeout 120 python tools/gpu_dbg_cjk.py cl100k_base 3000 2>&1 | grep -E "\[spl\]|match|counters|Error" | head -5
  done
done
