#!/bin/bash
# A/B of equi variants: prints jodo_equi avg ms and ms/step for each env setting
for v in "JODO_X=0" "JODO_EQUI_PAIR=1"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
  python - "$v" <<PY
import json,sys
d=json.load(open('/tmp/ab.json')); print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'equi', d['kernels']['jodo_equi']['avg_ms'], 'attn', d['kernels']['jodo_attn']['avg_ms'])
PY
done
