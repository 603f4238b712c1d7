#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -x -q -k "wide_mesh or energy" > $OUT/pytest_pk.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_pk.log; tail -n 3 $OUT/pytest_pk.log
run() { # tag env...
  TAG=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_pk_$TAG.json 2> $OUT/bench_pk_$TAG.err
  python -c "import json; d=json.load(open('$OUT/bench_pk_$TAG.json')); print('$TAG', round(d['ms_per_step'],3), round(d['roofline']['phase_ms_per_step']['PushDeposit'],3), round(d['roofline']['frac'],3))" || tail -n 3 $OUT/bench_pk_$TAG.err
}
run k5 EB200_PD_KERNEL=5
run k7 EB200_PD_KERNEL=7
