#!/bin/bash
# One GPU call: bench line, ncu launch list of the SAME bench command, ncu --set full of the hot
# kernels, and the march time-vs-levels sweep.  Outputs land in gpurun_out/ with the tag $1.
TAG=${1:-r1x}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cut -c1-400 gpurun_out/bench_${TAG}.json
python scripts/march_scaling.py > gpurun_out/march_scaling_${TAG}.jsonl 2>&1
cat gpurun_out/march_scaling_${TAG}.jsonl | cut -c1-300
# launch list of the bench command (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# full capture: 3rd march launch and the two back-transform passes after it
ncu --set full --clock-control none --import-source on -k regex:'k_march|k_fft' -s 6 -c 3 -f \
    -o gpurun_out/hot_${TAG} python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
