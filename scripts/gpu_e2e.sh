#!/bin/bash
# 1-GPU pass: parity tests, full bench line (streamed e2e), the same with the plain host step
TAG=${1:-r1n}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 5 $OUT/pytest_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; tail -n 3 $OUT/bench_$TAG.err
python -c "import json; d=json.load(open('$OUT/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'])"
EB200_HOST_STEP_PLAIN=1 timeout 900 python bench.py --steps 5 --no-cpu > $OUT/bench_${TAG}_plain.json 2> $OUT/bench_${TAG}_plain.err
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_plain.json')); print('plain', d['e2e'])"
for c in 2097152 33554432; do
EB200_HOST_CHUNK=$c timeout 900 python bench.py --steps 5 --no-cpu > $OUT/bench_${TAG}_c$c.json 2> $OUT/bench_${TAG}_c$c.err
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_c$c.json')); print('chunk $c', d['e2e'])"
done
