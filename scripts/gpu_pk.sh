#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { # tag env...
  TAG=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_pk_$TAG.json 2> $OUT/bench_pk_$TAG.err
  python -c "import json; d=json.load(open('$OUT/bench_pk_$TAG.json')); print('$TAG', round(d['ms_per_step'],3), round(d['roofline']['phase_ms_per_step']['PushDeposit'],3), round(d['roofline']['frac'],3))" || tail -n 3 $OUT/bench_pk_$TAG.err
}
run t256 A=1
run t128 EB200_VEC_THREADS=128
run t64 EB200_VEC_THREADS=64
