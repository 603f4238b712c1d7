#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_q.json 2> $OUT/bench_q.err; echo "rc=$?"
python -c "import json; d=json.load(open('$OUT/bench_q.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
