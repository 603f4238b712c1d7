#!/bin/bash
TAG=${1:-r1q_k6}
OUT=gpurun_out; mkdir -p $OUT
EB200_PD_KERNEL=6 EB200_PIPE_AHEAD=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_deposit_pipe -s 6 -c 1 \
    -f -o $OUT/prof_pd_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page details > $OUT/ncu_details_$TAG.txt 2>&1
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page source --csv > $OUT/ncu_source_$TAG.csv 2>&1
rm -f $OUT/prof_pd_$TAG.ncu-rep
