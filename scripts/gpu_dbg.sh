#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide_mesh or push_deposit_fast" 2>&1 | tail -25
for K in 3 4; do
  EB200_PD_KERNEL=$K timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_dbg_k$K.json 2> $OUT/bench_dbg_k$K.err
  echo "kernel $K: $(python -c "import json; d=json.load(open('$OUT/bench_dbg_k$K.json')); print(d['value']/1e9, d['roofline']['phase_ms_per_step']['PushDeposit'], d['roofline']['frac'])")"
done
