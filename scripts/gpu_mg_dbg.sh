#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
EB200_DECOMP2D=-1,2 EB200_DECOMP3D=-1,-1,-1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > $OUT/mg_worker.log 2>&1
echo "rc=$?"; grep -v "^W\|OMP_NUM\|\*\*\*\*" $OUT/mg_worker.log | tail -40
