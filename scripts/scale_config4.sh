#!/bin/bash
# BASELINE config 4 (8 towers x 1440 met steps) at 1/2/4/8 GPUs of one box
mkdir -p gpurun_out
P=29700
for G in 1 2 4 8; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P \
      scripts/bench_config4.py --steps 1440 2>&1 | grep '^{' | tee -a gpurun_out/config4_scale.jsonl
done
