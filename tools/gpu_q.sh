#!/bin/bash
# quick check: fused-path parity cases + QM9 (and optionally GEOM) bench kernels.  usage: bash tools/gpu_q.sh [geom]
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "qm9 or geom_l8 or geom_l10 or uniform" 2>&1 | tail -3
for w in qm9 $1; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > /tmp/q.json 2>/tmp/q.err || tail -3 /tmp/q.err
python - $w <<PY
import json,sys
d=json.load(open('/tmp/q.json')); k=d['kernels']
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), ' '.join(f"{n.replace('jodo_','')}={k[n]['avg_ms']}" for n in list(k)[:6]))
PY
done
