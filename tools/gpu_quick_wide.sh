#!/bin/bash
# Wide-path visit: all GPU tests + the nf=384 bench line (+ optionally other workloads: args after the tag)
OUT=gpurun_out/${1:-q3}; mkdir -p $OUT; shift
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for w in geom_large "$@"; do
timeout 600 python bench.py --workload $w --steps 6 --warmup 4 --no-cpu-baseline --no-extras > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"; tail -2 $OUT/bench_$w.err
python - <<PY
import json
d=json.load(open('$OUT/bench_$w.json'))
print('$w ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'tensor_frac', round(d['whole_step']['tensor_frac'],4))
for k,v in list(d['kernels'].items())[:22]: print('   ', k, v)
PY
done
