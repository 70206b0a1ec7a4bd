OUT=gpurun_out/s3q; mkdir -p $OUT
timeout 900 python -m pytest tests/test_sampler.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -3 $OUT/bench_qm9.err
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', d['e2e'])
PY
