#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_mg7.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/pytest_mg7.log
EB200_DECOMP2D=-1,2 EB200_DECOMP3D=-1,-1,-1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > $OUT/mg_worker.log 2>&1
echo "worker rc=$?"; grep "parity ok\|FAIL\|Error" $OUT/mg_worker.log | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_mg7_n$N.json 2> $OUT/bench_mg7_n$N.err
echo "N=$N: $(python -c "import json; d=json.loads(open('$OUT/bench_mg7_n$N.json').read().strip().split(chr(10))[-1]); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['parallelism'], d['roofline']['phase_ms_per_step'])")"
