#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list, ncu --set full of the named kernels.
# usage: tools/gpu_round.sh <tag> [kernel-regex ...]
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/${tag}_profile.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${tag}_full_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full_$k.log 2>&1
done
ls -la gpurun_out | tail -20
