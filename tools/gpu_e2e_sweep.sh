#!/bin/bash
# e2e of the bench workload against the pipeline's chunking parameters.  Usage: bash tools/gpu_e2e_sweep.sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-headline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  e2e %.2f GB/s  %.3f ms  (device %.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['device_ms_per_step']))"
}
for R in "4,2" "8,4,2" "16,8,4,2" "32,16,8,4,2"; do for C in 8 12 16; do
  echo "SPL_RAMP=$R SPL_CHUNKS_PER_DEV=$C"; SPL_RAMP=$R SPL_CHUNKS_PER_DEV=$C run
done; done
echo "--- trace (default)"; SPL_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-headline 2>&1 | grep "spl trace" | tail -14
