OUT=gpurun_out/s3i; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_imglinear.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
ONLY='c0 dots' timeout 300 python tools/bench_gemm_wide.py 980480 2>&1 | tee $OUT/gemm_new.txt
