#!/bin/bash
# Quick GPU visit: GPU tests + one bench line per workload.  usage: bash tools/gpu_quick.sh tag [pytest-args]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
shift
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench qm9 rc=$?"; tail -3 $OUT/bench_qm9.err
timeout 600 python bench.py --workload geom --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_geom.json 2> $OUT/bench_geom.err; echo "bench geom rc=$?"; tail -3 $OUT/bench_geom.err
timeout 600 python bench.py --workload geom_large --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_geom_large.json 2> $OUT/bench_geom_large.err; echo "bench geom_large rc=$?"; tail -3 $OUT/bench_geom_large.err
python - <<PY
import json
for w in ['qm9','geom','geom_large']:
    try:
        d=json.load(open('$OUT/bench_%s.json'%w))
        print(w, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
        for k,v in d['kernels'].items(): print('   ', k, v)
    except Exception as e: print(w, 'failed', e)
PY
