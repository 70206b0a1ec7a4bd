#!/bin/bash
# Same-box A/B of an environment switch on the headline bench: bash tools/gpu_ab_env.sh VAR [workload]
# prints ms/step and the top kernels with the switch unset and set to 1
V=$1; W=${2:-qm9}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for val in "" 1; do
  if [ -n "$val" ]; then export $V=$val; else unset $V; fi
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > /tmp/ab.json 2> /tmp/ab.err || tail -n 3 /tmp/ab.err
  python - "$V=$val" <<PY
import json, sys
d = json.load(open('/tmp/ab.json'))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), {k: v['avg_ms'] for k, v in list(d['kernels'].items())[:4]})
PY
done
