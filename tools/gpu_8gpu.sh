#!/bin/bash
# BASELINE configs[3] as quoted (GEOM nf=384, batch 4096 dealt over 8 B200s) + the QM9 headline on N GPUs.
N=${2:-8}
OUT=gpurun_out/${1:-g8}; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --workload geom_large --steps 8 --warmup 4 > $OUT/bench_geom_large_${N}gpu.json 2> $OUT/bench_geom_large_${N}gpu.err; echo "geom_large rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_qm9_${N}gpu.json 2> $OUT/bench_qm9_${N}gpu.err; echo "qm9 rc=$?"
python - <<PY
import json
for w in ('geom_large', 'qm9'):
    try:
        d = json.load(open('$OUT/bench_%s_${N}gpu.json' % w))
        print(w, 'n_gpus', d['n_gpus'], 'global batch', d['config']['global_batch'], 'ms/step', round(d['ms_per_step'], 3), 'value', round(d['value']),
              'e2e', round(d['e2e']['value']), 'rank min/max', round(d['rank_ms_per_step_min'], 3), round(d['rank_ms_per_step_max'], 3),
              'gather_ms', d['gather_ms'], 'strong', d['strong'] and {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['strong'].items()})
    except Exception as e:
        print(w, 'failed', e)
PY
for f in $OUT/*.err; do tail -n 2 $f; done
