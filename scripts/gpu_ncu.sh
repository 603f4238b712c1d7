#!/bin/bash
# launch list + one full ncu capture of the dominant kernel (default selection), summaries
TAG=${1:-r1p}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_deposit_vec -s 6 -c 1 \
    -f -o $OUT/prof_pd_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page details > $OUT/ncu_details_$TAG.txt 2>&1
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page source --csv > $OUT/ncu_source_$TAG.csv 2>&1
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > $OUT/ncu_dram_$TAG.csv 2>&1
ls -la $OUT | grep $TAG
