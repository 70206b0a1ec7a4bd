#!/bin/bash
# Full round-2 visit on the final code: GPU tests, headline bench (+ the other workloads inside its line), reference arm,
# separate GEOM lines, ncu launch lists, ncu full captures of the top kernels of both paths, sanitizers.
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench qm9 rc=$?"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload geom --steps 20 --warmup 4 --no-extras > $OUT/bench_geom.json 2> $OUT/bench_geom.err; echo "bench geom rc=$?"
timeout 600 python bench.py --workload geom_large --steps 10 --warmup 4 --no-extras > $OUT/bench_geom_large.json 2> $OUT/bench_geom_large.err; echo "bench geom_large rc=$?"
timeout 600 python bench.py --workload qm9_cond --steps 20 --warmup 4 --no-extras > $OUT/bench_qm9_cond.json 2> $OUT/bench_qm9_cond.err; echo "bench qm9_cond rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --noise philox --no-extras --no-cpu-baseline > $OUT/bench_qm9_philox.json 2> $OUT/bench_qm9_philox.err; echo "bench philox rc=$?"
JODO_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_run.log 2>&1; echo "ncu launches rc=$?"
JODO_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_geom_large.csv \
    python bench.py --workload geom_large --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_gl_run.log 2>&1; echo "ncu launches geom_large rc=$?"
for k in k_attn k_equi k_edge_update; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $OUT/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_$k.log 2>&1; echo "ncu $k rc=$?"
done
for k in k_wide_ln k_wide_attn_mol k_imglinear_dot2 k_wide_ffn_stream; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/prof_$k \
      python bench.py --workload geom_large --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_$k.log 2>&1; echo "ncu $k rc=$?"
done
bash tools/sanitize.sh $OUT > $OUT/sanitize_run.log 2>&1; echo "sanitize rc=$?"; tail -12 $OUT/sanitize_run.log
python - <<PY
import json
d=json.load(open('$OUT/bench_qm9.json'))
print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],1), 'launches', d['gpu_launches'])
print(json.dumps(d['roofline']), json.dumps(d['whole_step']))
print(json.dumps(d['workloads'], indent=1))
PY
