#!/bin/bash
# Evidence round: bench lines (own arm, reference arm), ncu launch list, ncu --set full of the encode kernels,
# per-config kernel times, decode + SentencePiece kernel captures, sanitizers.  Usage: bash tools/gpu_round2.sh <tag>
TAG=${1:-r01d}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${TAG}.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 python tools/gpu_cfgs.py > gpurun_out/cfgs_${TAG}.txt 2>&1
tail -8 gpurun_out/cfgs_${TAG}.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_probe|k_emit|k_pretok_fast|k_bpe' -s 12 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dec_|k_sp_|k_jl_' -c 40 -f -o gpurun_out/prof_aux_${TAG} \
    python tools/gpu_aux_kernels.py > gpurun_out/ncu_aux_${TAG}.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/sanitize_${TAG}.log 2>&1
tail -4 gpurun_out/sanitize_${TAG}.log
ls -la gpurun_out
