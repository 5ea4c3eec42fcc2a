#!/bin/bash
# bench under torchrun on N GPUs (+ the reference arm).  Usage: bash tools/gpu_multi.sh <tag> <N> [steps] [warmup]
TAG=$1; N=$2; K=${3:-10}; W=${4:-3}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${TAG}.log 2>&1
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu_${TAG}.txt 2>&1; numactl -H >> gpurun_out/lscpu_${TAG}.txt 2>&1
for g in $(seq 0 $((N-1))); do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | sed 's/^0000//' | tr A-Z a-z)/numa_node 2>/dev/null; done > gpurun_out/numa_${TAG}.txt 2>&1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps $K --warmup $W > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $K --warmup $W > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
fi
tail -c 2500 gpurun_out/bench_${TAG}.json; echo; tail -5 gpurun_out/bench_${TAG}.err
