#!/bin/bash
# multi-GPU pass (run under `gpurun --gpus N`): NCCL exchange tests + the N-rank bench line
N=${1:-2}; TAG=${2:-mg}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/smi_${TAG}.txt
timeout 1500 python -m pytest tests/test_gpu_metadomain.py -m gpu -q 2>&1 | tail -15 > $OUT/pytest_${TAG}.log
cat $OUT/pytest_${TAG}.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  fi
  echo "n=$n rc=$?"; tail -c 1500 $OUT/bench_${TAG}_n$n.json; tail -5 $OUT/bench_${TAG}_n$n.err
done
