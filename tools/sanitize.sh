#!/bin/bash
# compute-sanitizer over one small denoiser call (smoke): memcheck, racecheck (shared-memory hazards), synccheck.
OUT=gpurun_out/sanitize
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --report-api-errors no --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" $OUT/$tool.log | tail -3
done
# the nf = 384 wide path (row kernels of csrc/wide.cu + the GEMM)
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --report-api-errors no --print-limit 20 python -c "import __graft_entry__ as g; g.smoke('geom_large')" > $OUT/wide_$tool.log 2>&1
  echo "wide $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" $OUT/wide_$tool.log | tail -3
done
# the property classifier's row kernels and the in-kernel Philox update (memcheck over their GPU unit tests)
timeout 1200 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 20 python -m pytest tests/test_classifier.py tests/test_philox.py -m gpu -q \
    -k "fixture or oracle_draws or kernel_normals" > $OUT/extra_memcheck.log 2>&1
echo "extra memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/extra_memcheck.log | tail -3
# the fused wide-path edge FFN in every compiled size, the row-0 matrix-vector kernel, the dots-only GEMM (unit tests)
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --report-api-errors no --print-limit 20 python -m pytest tests/test_gpu_wide_units.py tests/test_gpu_rowlinear.py \
      tests/test_gpu_imglinear.py -m gpu -q -k "(fused_edge_ffn and not 700) or row0_linear or row_dots" > $OUT/units_$tool.log 2>&1
  echo "units $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/units_$tool.log | tail -3
done
