#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list, ncu full captures of the edge kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "bench qm9 rc=$?"
timeout 600 python bench.py --workload geom --steps 20 --warmup 3 > $OUT/bench_geom.json 2> $OUT/bench_geom.err; echo "bench geom rc=$?"
timeout 600 python bench.py --workload geom_l10 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_geom_l10.json 2> $OUT/bench_geom_l10.err; echo "bench geom_l10 rc=$?"
timeout 600 python bench.py --workload geom_large --steps 10 --warmup 3 > $OUT/bench_geom_large.json 2> $OUT/bench_geom_large.err; echo "bench geom_large rc=$?"
timeout 600 python bench.py --workload qm9_cond --steps 20 --warmup 4 > $OUT/bench_qm9_cond.json 2> $OUT/bench_qm9_cond.err; echo "bench qm9_cond rc=$?"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
kill $SMI
JODO_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_run.log 2>&1; echo "ncu launches rc=$?"
for k in k_attn k_equi k_edge_update k_imglinear; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 2 -f -o $OUT/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/prof_$k.log 2>&1; echo "ncu $k rc=$?"
done
for k in k_wide_ln k_wide_attn; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/prof_$k \
      python bench.py --workload geom_large --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/prof_$k.log 2>&1; echo "ncu $k rc=$?"
done
cat $OUT/bench_qm9.json | head -c 3000
