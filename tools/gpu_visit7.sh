#!/bin/bash
OUT=gpurun_out/${1:-v7}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|chain, 8|sampled molecules|clamped|config0|^E  " $OUT/pytest_gpu.log | tail -24
