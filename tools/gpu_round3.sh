#!/bin/bash
# Evidence round (k_bpe / k_bpe_long split): GPU tests, bench lines (own arm, reference arm), per-config kernel times,
# ncu launch list, ncu --set full of the encode kernels and of the merge kernels on cfg4 / cfg5.
# Usage: bash tools/gpu_round3.sh <tag>
TAG=${1:-r01f}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1
tail -8 gpurun_out/cfgs_${TAG}.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_probe|k_emit|k_pretok_fast|k_bpe' -s 15 -c 5 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
bash tools/gpu_ncu_bpe.sh ${TAG}
ls -la gpurun_out | tail -30
