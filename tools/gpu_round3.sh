#!/bin/bash
# usage: tools/gpu_round3.sh <tag> [ncu-kernel-regex]   -- tests, bench (+ family profile), torch-profiler timeline,
# ncu launch list, optional ncu --set full of the kernels matching the regex (1-layer step; raw CSV only, the
# .ncu-rep stays on the box: gpurun_out/ must stay under 64 MiB)
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
python bench.py --no-cpu-baseline --profile-out gpurun_out/${tag}_profile.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
python tools/profile_step.py > gpurun_out/${tag}_timeline.txt 2>&1; head -30 gpurun_out/${tag}_timeline.txt
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -f -o /tmp/${tag}_full \
    python tools/profile_once.py --layers 1 > gpurun_out/${tag}_ncu_full.log 2>&1
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
  ls -la /tmp/${tag}_full.ncu-rep gpurun_out/${tag}_full_raw.csv
fi
du -sh gpurun_out
