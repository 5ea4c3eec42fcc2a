#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -k "pipelined or multi_device or cfg1 or cfg2 or special or golden or jsonl or sentencepiece or invalid or huge" 2>&1 | tail -4
for G in 1 0; do
  echo "== SPL_GRAPH=$G"; SPL_GRAPH=$G timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph$G.json 2> gpurun_out/bench_graph$G.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_graph$G.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],2),'ms',round(d['e2e']['ms_per_step'],3),'dev',round(d['e2e']['device_ms_per_step'],3),'ceil',round(d['e2e']['pcie_ceiling']['value'],1))
for c,v in d['configs'].items(): print(' ',c,round(v['value'],1),'e2e',round(v['e2e']['value'],2),'ms',round(v['e2e']['ms_per_step'],3))
print(' small', {k:(round(v,1) if isinstance(v,float) else v) for k,v in (d.get('small_batch') or {}).items() if k!='note'})
PY
done
