#!/bin/bash
# Evidence for profiles/: ncu launch list of the bench step, ncu --set full of every kernel of one step on cfg2 (the bench
# workload), cfg4 and cfg5.  Usage: bash tools/gpu_evidence.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${TAG}.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-headline > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
K='k_mark_docs|k_pretok_fast|k_pretok_fb|k_probe|k_bpe|k_emit'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 24 -c 8 -f -o gpurun_out/prof_cfg2_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-headline > gpurun_out/ncu_cfg2_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_cfg2_${TAG}.log | cut -c1-200
for c in cfg4 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 16 -c 8 -f -o gpurun_out/prof_${c}_${TAG} \
      python tools/gpu_cfgs.py $c > gpurun_out/ncu_${c}_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_${c}_${TAG}.log | cut -c1-200
done
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1
SPL_NO_DEDUP=1 timeout 600 python tools/gpu_cfgs.py cfg4 > gpurun_out/cfgs_${TAG}_nodedup.txt 2>&1
SPL_PROBE_BULK=0 timeout 600 python tools/gpu_cfgs.py cfg2 > gpurun_out/cfgs_${TAG}_nobulk.txt 2>&1
cat gpurun_out/cfgs_${TAG}.txt | cut -c1-250
ls -la gpurun_out/*${TAG}* | tail
