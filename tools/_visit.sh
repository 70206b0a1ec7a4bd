OUT=gpurun_out/s3a; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_imglinear.py tests/test_gpu_parity.py tests/test_gpu_wide_units.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
ONLY='c0 dots' timeout 300 python tools/bench_gemm_wide.py 980480 2>&1 | tee $OUT/gemm_new.txt
JODO_DOT_LEGACY=1 ONLY='c0 dots' timeout 300 python tools/bench_gemm_wide.py 980480 2>&1 | tee $OUT/gemm_legacy.txt
timeout 600 python bench.py --workload geom_large --steps 6 --warmup 4 --no-cpu-baseline --no-extras > $OUT/bench_gl.json 2> $OUT/bench_gl.err; echo "bench rc=$?"; tail -2 $OUT/bench_gl.err
python - <<PY
import json
d=json.load(open('$OUT/bench_gl.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
for k,v in list(d['kernels'].items())[:8]: print('   ', k, v)
PY
