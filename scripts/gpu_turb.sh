#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { TAG=$1; shift
env $ENVV timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > $OUT/bench_turb_$TAG.json 2> $OUT/bench_turb_$TAG.err; echo "rc=$?"; tail -n 3 $OUT/bench_turb_$TAG.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_turb_$TAG.json')); print('$TAG', round(d['value']/1e9,3), round(d['ms_per_step'],1), d['config']['sort_interval'], round(d['roofline']['phase_ms_per_step']['PushDeposit'],1))"
}
ENVV="EB200_PD_KERNEL=1" run k1_agg --turbulence 160 --deposit aggregated
ENVV="EB200_PD_KERNEL=1" run k1_agg_s2 --turbulence 160 --deposit aggregated --sort-interval 2
ENVV="A=1" run auto --turbulence 160
