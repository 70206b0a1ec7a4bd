#!/bin/bash
# ncu --set full capture of one kernel of the headline bench: bash tools/gpu_prof1.sh tag kernel-regex [workload]
OUT=gpurun_out/${1:-p1}; mkdir -p $OUT
K=${2:-k_equi_lin}
W=${3:-qm9}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 1 -f -o $OUT/prof_$K \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_$K.log 2>&1; echo "ncu $K rc=$?"
