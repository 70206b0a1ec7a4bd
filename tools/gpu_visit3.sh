#!/bin/bash
OUT=gpurun_out/${1:-v3}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_reference_dropin.py -m gpu -q -s 2>&1 | grep -v Warning | tail -6
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
python - <<PY
import json
for w in ['qm9']:
    try:
        d=json.load(open('$OUT/bench_%s.json'%w))
        print(w, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d.get('cpu_baseline'))
        print(d['roofline'], d['whole_step'])
        for k,v in d['kernels'].items(): print('   ', k, v)
    except Exception as e: print(w, 'failed', e)
PY
