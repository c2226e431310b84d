#!/bin/bash
# usage: tools/ncu_attn.sh <tag>   -- ncu --set full of one forward and one backward launch of the config-5 attention core
tag=$1
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:attn_reg" --launch-skip 2 --launch-count 2 -f -o /tmp/${tag}_a \
  python tools/attn_wide.py 0.1 > gpurun_out/${tag}.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:attn_reg_bwd" --launch-skip 2 --launch-count 1 -f -o /tmp/${tag}_b \
  python tools/attn_wide.py 0.1 >> gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
ncu -i /tmp/${tag}_a.ncu-rep --page raw --csv > gpurun_out/${tag}_fwd.raw.csv 2>/dev/null
ncu -i /tmp/${tag}_b.ncu-rep --page raw --csv > gpurun_out/${tag}_bwd.raw.csv 2>/dev/null
ncu -i /tmp/${tag}_a.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_fwd.src.csv 2>/dev/null
ncu -i /tmp/${tag}_b.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_bwd.src.csv 2>/dev/null
ls -la gpurun_out/${tag}*
