#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
EB200_DECOMP2D=-1,2 EB200_DECOMP3D=-1,-1,-1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > $OUT/mg_worker.log 2>&1
echo "worker rc=$?"; grep "parity ok\|FAIL\|Error" $OUT/mg_worker.log | head
bash scripts/gpu_mg4.sh $N
