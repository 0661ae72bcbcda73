#!/bin/bash
# Config 5 (4096^2 x 414 levels, one solve) ky-slab sharded over the G ranks given as arguments.
mkdir -p gpurun_out
P=29700
for G in "$@"; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P \
      scripts/bench_sharded.py --reps 4 $( [ $G -eq 2 ] && echo --check ) 2>&1 | grep '^{' | tee -a gpurun_out/sharded_scale_r1h.jsonl
done
