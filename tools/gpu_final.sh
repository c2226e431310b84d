#!/bin/bash
# Round-end evidence in one gpurun call: GPU tests, smoke(), bench line (1M headline + TG block + CPU / eager baselines
# + per-family profile), the reference arm, the ncu launch list of the bench command, and one ncu --set full capture of
# every hot kernel of a 2-layer step on the 1M graph (raw CSV; the .ncu-rep stays on the box: 64 MiB return limit).
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --profile-out gpurun_out/${tag}_profile_1M.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
head -c 400 gpurun_out/${tag}_bench.json; echo
python bench.py --workload TG --no-cpu-baseline --no-gpu-baseline --steps 50 --warmup 8 --profile-out gpurun_out/${tag}_profile_TG.json > gpurun_out/${tag}_bench_TG.json 2> gpurun_out/${tag}_bench_TG.err
head -c 300 gpurun_out/${tag}_bench_TG.json; echo
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
head -c 300 gpurun_out/${tag}_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-secondary > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k "regex:dw_tile_kernel|ln_bwd_stream|sample_contexts|linear_tile_kernel|attn_mma|embed_|gather_proj" -f -o /tmp/${tag}_full \
  python tools/profile_once.py --layers 2 --workload 1M > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | grep ${tag}
