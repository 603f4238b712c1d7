#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_stats.py -m gpu -x -q > $OUT/pytest_stats.log 2>&1; echo "rc=$?"; tail -n 15 $OUT/pytest_stats.log
