#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_last.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/pytest_last.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
