#!/bin/bash
# round-2 visit 1: new boundary / quirk tests, full gpu suite, baseline bench + reference arm, MUFU microbench
OUT=gpurun_out/v1; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mufu_rate tools/micro/mufu_rate.cu && tools/micro/mufu_rate > $OUT/mufu.txt 2>&1; cat $OUT/mufu.txt
timeout 900 python -m pytest tests/test_reference_dropin.py tests/test_gpu_quirks.py -m gpu -q -s > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -25 $OUT/pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_quirks.py > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json | head -c 600
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['cpu_baseline'])
for k,v in d['kernels'].items(): print('   ', k, v)
PY
