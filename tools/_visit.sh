OUT=gpurun_out/s3o; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide_units.py tests/test_gpu_quirks.py tests/test_sampler.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
show() { python - <<PY
import json
d=json.load(open('$1'))
print('$1 ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:$2]: print('   ', k, v['launches_per_step'], v['avg_ms'])
PY
}
timeout 600 python bench.py --workload geom_large --steps 6 --warmup 4 --no-cpu-baseline --no-extras > $OUT/bench_gl.json 2> $OUT/bench_gl.err; echo "bench rc=$?"; tail -2 $OUT/bench_gl.err; show $OUT/bench_gl.json 18
