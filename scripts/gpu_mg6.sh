#!/bin/bash
# 2-GPU migration diagnosis: x-split (2x1) vs y-split (1x2), stage trace of the migration
OUT=gpurun_out; mkdir -p $OUT
run() { # tag decomp0 decomp1 steps
  TAG=$1
  EB200_COMM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps $4 --warmup 3 --no-e2e --no-cpu --decomp $2 $3 > $OUT/bench_mg6_$TAG.json 2> $OUT/bench_mg6_$TAG.err
  echo "$TAG: $(python -c "import json; d=json.loads(open('$OUT/bench_mg6_$TAG.json').read().strip().split(chr(10))[-1]); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['parallelism'], d['roofline']['phase_ms_per_step'])")"
  grep migrate $OUT/bench_mg6_$TAG.err | tail -4
}
run x 2 1 20
run y 1 2 20
