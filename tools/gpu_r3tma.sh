#!/bin/bash
# Evidence run after the TMA row gather became the default of gather_proj.cu: GPU tests, smoke(), bench line (1M headline
# + TG block + per-family profile; the CPU / eager baselines are unchanged and skipped here).
tag=$1
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log; tail -2 gpurun_out/${tag}_smoke.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/${tag}_profile_1M.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
head -c 300 gpurun_out/${tag}_bench.json; echo
