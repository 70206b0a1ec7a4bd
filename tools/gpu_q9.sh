#!/bin/bash
OUT=gpurun_out/${1:-q9}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_imglinear.py tests/test_gpu_rowlinear.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
for w in qm9 geom_large; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -n 2 $OUT/bench_$w.err
python - <<PY
import json
d=json.load(open('$OUT/bench_$w.json'))
print('$w ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
for k,v in list(d['kernels'].items())[:14]: print('   ', k, v)
PY
done
