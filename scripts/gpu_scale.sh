#!/bin/bash
# weak-scaling bench lines at N ranks (run under `gpurun --gpus N`)
N=${1:-8}; TAG=${2:-scale}
OUT=gpurun_out; mkdir -p $OUT
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $n --steps 40 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  fi
  echo "n=$n rc=$? $(python -c "import json; d=json.load(open('$OUT/bench_${TAG}_n$n.json')); print(round(d['value']/1e9,2), d['ms_per_step'], d['config']['parallelism'], d['roofline']['phase_ms_per_step'])")"
  tail -3 $OUT/bench_${TAG}_n$n.err | grep -v OMP_NUM
done
