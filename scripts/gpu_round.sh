#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, one full ncu capture of the top kernel.
# usage (under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"
cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_deposit_vec -s 6 -c 1 \
    -f -o $OUT/prof_pd_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page details > $OUT/ncu_details_$TAG.txt 2>&1
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page source --csv > $OUT/ncu_source_$TAG.csv 2>&1
ncu -i $OUT/prof_pd_$TAG.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > $OUT/ncu_dram_$TAG.csv 2>&1
