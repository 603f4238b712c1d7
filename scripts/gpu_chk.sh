#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for SI in 20 40; do
  timeout 600 python bench.py --steps 80 --warmup 3 --no-e2e --no-cpu --sort-interval $SI > $OUT/bench_chk_$SI.json 2> $OUT/bench_chk_$SI.err
  echo "sort-interval $SI: $(python -c "import json; d=json.load(open('$OUT/bench_chk_$SI.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['phase_ms_per_step'])")"
done
