#!/bin/bash
# Scaling runs on one 8-GPU box: config 5 sharded at G=1,2,4,8 and the headline bench at N=1,2,4,8.
mkdir -p gpurun_out
P=29600
for G in 1 2 4 8; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P \
      scripts/bench_sharded.py --reps 4 $( [ $G -eq 2 ] && echo --check ) 2>&1 | grep '^{' | tee -a gpurun_out/sharded_scale.jsonl
done
for G in 1 2 4 8; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $G --steps 200 --warmup 5 --no-cpu 2>&1 | grep '^{' | tee -a gpurun_out/bench_scale.jsonl | cut -c1-260
done
