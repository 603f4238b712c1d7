#!/bin/bash
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 40 --warmup 3 --no-cpu > $OUT/bench_r1r_n$N.json 2> $OUT/bench_r1r_n$N.err
echo "rc=$?"; tail -n 3 $OUT/bench_r1r_n$N.err | grep -v OMP
python -c "import json; d=json.load(open('$OUT/bench_r1r_n$N.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['parallelism'], d['roofline']['phase_ms_per_step'], d['e2e'])"
