#!/bin/bash
# One gpurun call: GPU tests, bench line + per-family profile, ncu launch list of the bench command,
# one ncu --set full capture of a 1-layer step (every kernel family once).
# usage: tools/gpu_round2.sh <tag> [skip-tests]
tag=$1
mkdir -p gpurun_out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -3 gpurun_out/${tag}_pytest.log
fi
python bench.py --profile-out gpurun_out/${tag}_profile.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${tag}_full \
  python tools/profile_once.py --layers 1 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out | tail -12
