#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_turb.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_turb.log
run() { TAG=$1; shift
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > $OUT/bench_turb_$TAG.json 2> $OUT/bench_turb_$TAG.err; echo "rc=$?"; tail -n 3 $OUT/bench_turb_$TAG.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_turb_$TAG.json')); print('$TAG', round(d['value']/1e9,3), round(d['ms_per_step'],1), d['config']['particles_per_gpu'], d['config']['sort_interval'], d['roofline']['phase_ms_per_step'])"
}
run n192 --turbulence 192
