#!/bin/bash
# Quick A/B visit: parity tests of the fused path + headline bench without extras (+ optional microbenchmarks)
OUT=gpurun_out/${1:-q2}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_precision.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:10]: print('   ', k, v)
PY
if [ -n "$2" ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate tools/micro/mufu_rate.cu && /tmp/mufu_rate | tee $OUT/mufu_rate.txt
fi
