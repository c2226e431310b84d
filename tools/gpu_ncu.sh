#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel-regex> [profile_once args...]
# ncu --set full of the matching kernels of ONE step (1-layer model unless overridden); brings back the raw-page CSV
# (the .ncu-rep stays on the box: gpurun_out/ must stay under 64 MiB).
tag=$1; rx=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off -k "regex:$rx" -f -o /tmp/${tag}_full \
  python tools/profile_once.py "$@" > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ls -la /tmp/${tag}_full.ncu-rep gpurun_out/${tag}_full_raw.csv
