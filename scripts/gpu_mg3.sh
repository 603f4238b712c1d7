#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { # tag steps env...
  TAG=$1; STEPS=$2; shift 2
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps $STEPS --warmup 3 --no-e2e --no-cpu > $OUT/bench_mg3_$TAG.json 2> $OUT/bench_mg3_$TAG.err
  echo "$TAG: $(wc -l < $OUT/bench_mg3_$TAG.json) stdout lines; $(python -c "import json; d=json.loads(open('$OUT/bench_mg3_$TAG.json').read().strip().split(chr(10))[-1]); print(round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['phase_ms_per_step'])")"
}
run s10 10 A=1
run s40 40 A=1
run s10_nopair 10 EB200_NO_FILTER_FUSION=1
run s40_nopair 40 EB200_NO_FILTER_FUSION=1
