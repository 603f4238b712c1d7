#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -q -x -k "filter or step" 2>&1 | tail -6
for SI in 10 20 40 80; do
  timeout 600 python bench.py --steps 80 --warmup 3 --no-e2e --no-cpu --sort-interval $SI > $OUT/bench_sort_$SI.json 2> $OUT/bench_sort_$SI.err
  echo "sort-interval $SI: $(python -c "import json; d=json.load(open('$OUT/bench_sort_$SI.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['phase_ms_per_step'])")"
done
