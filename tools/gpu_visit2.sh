#!/bin/bash
OUT=gpurun_out/${1:-v2}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
timeout 600 python bench.py --workload geom --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_geom.json 2> $OUT/bench_geom.err; echo "bench geom rc=$?"; tail -3 $OUT/bench_geom.err
python - <<PY
import json
for w in ['qm9','geom']:
    try:
        d=json.load(open('$OUT/bench_%s.json'%w))
        print(w, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d.get('cpu_baseline'))
        for k,v in d['kernels'].items(): print('   ', k, v)
    except Exception as e: print(w, 'failed', e)
PY
if [ "$2" == "san" ]; then
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/memcheck.log
fi
