#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --walls --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_walls.json 2> $OUT/bench_walls.err; echo "rc=$?"; tail -n 3 $OUT/bench_walls.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_walls.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['workload'][-70:], d['roofline']['phase_ms_per_step'])"
