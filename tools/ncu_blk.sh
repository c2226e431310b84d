#!/bin/bash
# usage: tools/ncu_blk.sh <tag> <which: fused_fwd,fused_bwd,dense_fwd,dense_bwd>
# ncu --set full (with source counters) of ONE launch of each requested fused-block kernel at the bench token count;
# brings back the raw-page and source-page CSVs.
tag=$1; which=$2
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:blk_" --launch-skip 3 --launch-count 1 -f -o /tmp/${tag} \
  python tools/bench_ffn.py 294912 0.1 $which > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}.raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}.src.csv 2>/dev/null
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}.*.csv
