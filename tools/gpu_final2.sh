#!/bin/bash
# final evidence pass: ncu launch list + --set full per config (tools/gpu_evidence.sh), the Parquet page kernel, sanitizers.
TAG=${1:-r02z}
bash tools/gpu_evidence.sh $TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pq_pages -c 1 -f -o gpurun_out/prof_pq_${TAG} \
    python tools/gpu_parquet.py "snappy, PLAIN" > gpurun_out/ncu_pq_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_pq_${TAG}.log | cut -c1-200
bash tools/gpu_sanitize.sh
