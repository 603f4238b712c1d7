#!/bin/bash
# Push+deposit kernel comparison on the GPU box: parity tests of the particle path, one short
# bench per fused kernel (EB200_PD_KERNEL = 2 TMA stream, 3 vec4), one full ncu capture.
# usage (under gpurun): bash scripts/gpu_pd.sh <tag> [kernel-regex]
TAG=${1:-pd}
KRE=${2:-push_deposit_vec}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
for K in 2 3; do
  EB200_PD_KERNEL=$K timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_k$K.json 2> $OUT/bench_${TAG}_k$K.err
  echo "kernel $K rc=$?"; cat $OUT/bench_${TAG}_k$K.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 6 -c 1 \
    -f -o $OUT/prof_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page details > $OUT/ncu_details_$TAG.txt 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/ncu_source_$TAG.csv 2>&1
ls -la $OUT
