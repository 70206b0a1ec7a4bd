OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 600 python bench.py --workload qm9_cond --steps 20 --warmup 4 --no-extras > $OUT/bench_qm9_cond.json 2> $OUT/bench_qm9_cond.err; echo "bench qm9_cond rc=$?"; tail -3 $OUT/bench_qm9_cond.err
python - <<PY
import json
for w in ('qm9_cond',):
    d=json.load(open('$OUT/bench_%s.json' % w))
    print(w, 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'in line', d['e2e'].get('in_line_value') and round(d['e2e']['in_line_value']), 'finite', d.get('finite'), d['e2e'])
PY
