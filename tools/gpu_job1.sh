#!/bin/bash
# one GPU-box visit: parity tests, bench line, ncu full captures (never a bench number)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
ls /root/repo/MEASURED_PEAKS.json && cp /root/repo/MEASURED_PEAKS.json gpurun_out/ 
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh2d -s 20 -c 2 -f -o gpurun_out/prof_mesh python tools/prof_target.py mesh > gpurun_out/ncu_mesh.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cols_fast|rows_fwd_fast|rows_inv_fast|peak2|patch_' -s 10 -c 5 -f -o gpurun_out/prof_flow python tools/prof_target.py flow > gpurun_out/ncu_flow.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-600
