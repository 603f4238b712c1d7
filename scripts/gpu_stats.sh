#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "match" > $OUT/pytest_stats.log 2>&1; echo "rc=$?"; tail -n 25 $OUT/pytest_stats.log
