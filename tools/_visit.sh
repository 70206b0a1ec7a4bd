OUT=gpurun_out/s3n; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_quirks.py tests/test_gpu_precision.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
show() { python - <<PY
import json
d=json.load(open('$1'))
print('$1 ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:$2]: print('   ', k, v['launches_per_step'], v['avg_ms'])
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err; show $OUT/bench_qm9.json 4
