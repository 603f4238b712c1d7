#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -x -q -k "filter or match or strict_exact" > $OUT/pytest_walls.log 2>&1; echo "pytest rc=$?"; tail -n 3 $OUT/pytest_walls.log
for w in "--walls" ""; do
timeout 600 python bench.py $w --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_walls.json 2> $OUT/bench_walls.err; echo "rc=$?"; tail -n 3 $OUT/bench_walls.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_walls.json')); print('$w', round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
[ -n "$w" ] && cp $OUT/bench_walls.json $OUT/bench_walls_w.json
done
