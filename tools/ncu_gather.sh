#!/bin/bash
# usage: tools/ncu_gather.sh <tag>   -- ncu --set full of one forward and one dW launch of the gather-fused projections
tag=$1
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:gather_proj" --launch-skip 12 --launch-count 2 -f -o /tmp/${tag} \
  python tools/bench_gather.py > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}.raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}.src.csv 2>/dev/null
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}.*.csv
