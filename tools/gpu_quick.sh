#!/bin/bash
# Quick confirmation of HEAD in one gpurun call: GPU tests, smoke(), the default bench line (no ncu).
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --profile-out gpurun_out/${tag}_profile_1M.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
head -c 300 gpurun_out/${tag}_bench.json; echo
