#!/bin/bash
# Round-2 GPU-box visit: parity tests, headline bench (+ other workloads inside its line), reference arm, ncu launch list,
# ncu full captures of the edge kernels.  usage (under gpurun): bash tools/gpu_round2.sh [tag] [skip-tests]
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench qm9 rc=$?"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
JODO_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_run.log 2>&1; echo "ncu launches rc=$?"
for k in k_attn k_equi k_edge_update; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $OUT/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_$k.log 2>&1; echo "ncu $k rc=$?"
done
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1), 'launches', d['gpu_launches'])
print(json.dumps(d['roofline']), json.dumps(d['whole_step']))
print(json.dumps(d['workloads'], indent=1))
for k,v in list(d['kernels'].items())[:12]: print(k, v)
PY
