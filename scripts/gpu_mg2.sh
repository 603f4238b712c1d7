#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter" 2>&1 | tail -3
EB200_DECOMP2D=-1,2 EB200_DECOMP3D=-1,-1,-1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > $OUT/mg_worker.log 2>&1
echo "worker rc=$?"; grep "parity ok\|FAIL" $OUT/mg_worker.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_mg2.json 2> $OUT/bench_mg2.err
echo "bench rc=$?"; cat $OUT/bench_mg2.json | head -c 3000
