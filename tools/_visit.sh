OUT=gpurun_out/s3m; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rowlinear.py tests/test_gpu_parity.py tests/test_gpu_quirks.py tests/test_gpu_boundary.py tests/test_sampler.py tests/test_gpu_precision.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
