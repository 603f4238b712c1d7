#!/bin/bash
# round-end style check on one GPU: parity tests, smoke(), default bench line, reference arm
TAG=${1:-r1u}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log; tail -n 3 $OUT/pytest_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; tail -n 3 $OUT/bench_$TAG.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cut -c1-400 $OUT/bench_ref_$TAG.json
