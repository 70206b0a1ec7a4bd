#!/bin/bash
OUT=gpurun_out/${1:-q6}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "uniform or qm9_selfcond or geom_l8" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench rc=$?"; tail -2 $OUT/bench_qm9.err
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
for k,v in list(d['kernels'].items())[:4]: print('   ', k, v)
PY
JODO_NVCC_EXTRA=-DJODO_PHASE_TIMING timeout 600 python tools/phase_timing.py qm9 2>&1 | tail -40 | tee $OUT/phases.txt
