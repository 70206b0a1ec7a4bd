#!/bin/bash
OUT=gpurun_out/${1:-v6}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|dpm chain|eager" | tail
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
timeout 300 python bench.py --workload qm9_cond --steps 20 --warmup 4 --no-cpu-baseline > $OUT/bench_qm9_cond.json 2> $OUT/bench_qm9_cond.err; echo "bench cond rc=$?"; tail -2 $OUT/bench_qm9_cond.err
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1), 'launches', d['gpu_launches'])
print(json.dumps(d['workloads'], indent=1))
d=json.load(open('$OUT/bench_qm9_cond.json'))
print('cond ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['step_launch'])
PY
