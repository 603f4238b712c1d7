#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_curvilinear.py -m gpu -q 2>&1 | tail -40 > $OUT/pytest_curv.log
cat $OUT/pytest_curv.log
