OUT=gpurun_out/s3r; mkdir -p $OUT
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_parity.py tests/test_gpu_imglinear.py tests/test_gpu_rowlinear.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
for v in pdl nopdl; do
  if [ $v = nopdl ]; then export JODO_NO_PDL=1; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_qm9_$v.json 2> $OUT/bench_qm9_$v.err; echo "bench $v rc=$?"; tail -2 $OUT/bench_qm9_$v.err
  python - <<PY
import json
d=json.load(open('$OUT/bench_qm9_$v.json'))
print('$v ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'in line', round(d['e2e']['in_line_value']))
PY
done
