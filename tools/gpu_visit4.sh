#!/bin/bash
OUT=gpurun_out/${1:-v4}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -12 $OUT/pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_sampler.py tests/test_gpu_quirks.py tests/test_gpu_boundary.py -m gpu -q > $OUT/pytest_more.log 2>&1; echo "more rc=$?"; tail -6 $OUT/pytest_more.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
JODO_EQUI_SINGLE=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_qm9_single.json 2> $OUT/bench_qm9_single.err; echo "bench single rc=$?"
timeout 600 python bench.py --workload geom --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_geom.json 2> $OUT/bench_geom.err; echo "bench geom rc=$?"
python - <<PY
import json
for w in ['qm9','qm9_single','geom']:
    try:
        d=json.load(open('$OUT/bench_%s.json'%w))
        print(w, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), 'whole', round(d['whole_step']['tensor_frac'],3))
        for k,v in list(d['kernels'].items())[:4]: print('   ', k, v)
    except Exception as e: print(w, 'failed', e)
PY
