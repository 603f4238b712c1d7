#!/bin/bash
# bench the fused push+deposit kernel variants built by entity_b200.build.build_variant
# usage (under gpurun): bash scripts/gpu_variants.sh <tag> <kernel> <variant> [<variant> ...]
TAG=$1; K=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for V in "$@"; do
  if [ "$V" = base ]; then unset EB200_LIB; else export EB200_LIB=$PWD/entity_b200/variants/$V.so; fi
  EB200_PD_KERNEL=$K timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_$V.json 2> $OUT/bench_${TAG}_$V.err
  echo "$V rc=$? $(python -c "import json,sys; d=json.load(open('$OUT/bench_${TAG}_$V.json')); print(d['value']/1e9, d['roofline']['phase_ms_per_step']['PushDeposit'], d['roofline']['frac'])")"
done
