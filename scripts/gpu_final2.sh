#!/bin/bash
TAG=${1:-r1y}
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_final.sh $TAG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
