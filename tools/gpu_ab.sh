#!/bin/bash
# A/B of an environment switch on the bench (kernel table only).  Usage: bash tools/gpu_ab.sh VAR [steps]
VAR=$1; STEPS=${2:-20}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for v in 0 1; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value %.1f GB/s  e2e %.1f GB/s  ms/step %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print({k: round(v*1000,1) for k,v in d['roofline']['kernel_ms'].items()})
"
done
