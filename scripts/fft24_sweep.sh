#!/bin/bash
# launch-shape sweep of the specialised back-transform passes (transforms per CTA x threads per CTA)
for n in 512 1024; do
  if [ $n -eq 512 ]; then ARGS="--n 512"; else ARGS="--n 1024 --nz 128 --groups 2 --towers 16"; fi
  for cx in 1 2 4; do for tx in 192 256 384; do
    echo "n=$n X cw=$cx thr=$tx $(BLDFM_FFT24_CW_X=$cx BLDFM_FFT24_THREADS_X=$tx python scripts/fft_throughput.py $ARGS --reps 3 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["inverse_ms"],4))')"
  done; done
  for cy in 1 2 4; do for ty in 192 256 384; do
    echo "n=$n Y cw=$cy thr=$ty $(BLDFM_FFT24_CW_Y=$cy BLDFM_FFT24_THREADS_Y=$ty python scripts/fft_throughput.py $ARGS --reps 3 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["inverse_ms"],4))')"
  done; done
done
