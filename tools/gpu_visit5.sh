#!/bin/bash
OUT=gpurun_out/${1:-v5}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|packed in|repack|config0" $OUT/pytest_gpu.log | tail -20
timeout 300 python -c "
import torch, __graft_entry__ as g
from jodo_b200 import _lib
g.smoke(); print('launches through the C ABI in smoke():', _lib.LAUNCHES)"
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1))
PY
