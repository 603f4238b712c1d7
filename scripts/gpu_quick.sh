#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 80 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick.json 2> $OUT/bench_quick.err
echo "bench: $(python -c "import json; d=json.load(open('$OUT/bench_quick.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'])")"
