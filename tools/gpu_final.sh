#!/bin/bash
# Round-end evidence in one gpurun call: GPU tests, smoke(), bench line (+ CPU baseline + per-family profile), the
# reference arm, the ncu launch list of the bench command, and one ncu --set full capture of every hot kernel of a
# 2-layer step (raw CSV; the .ncu-rep is kept only if it is small enough for the 64 MiB return limit).
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --profile-out gpurun_out/${tag}_profile.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cat gpurun_out/${tag}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k "regex:dw_tile_kernel|ln_bwd_stream|sample_contexts|linear_tile_kernel|attn_mma|embed_" -f -o /tmp/${tag}_full \
  python tools/profile_once.py --layers 2 > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/${tag}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/${tag}_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out; ls -la gpurun_out | tail -15
