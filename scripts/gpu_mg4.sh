#!/bin/bash
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_mg4_n$N.json 2> $OUT/bench_mg4_n$N.err
echo "N=$N: $(python -c "import json; d=json.loads(open('$OUT/bench_mg4_n$N.json').read().strip().split(chr(10))[-1]); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['parallelism'], d['roofline']['phase_ms_per_step'])")"
