#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --path mesh --no-cpu-baseline > gpurun_out/bench_mesh.json 2> gpurun_out/bench_mesh.err
cat gpurun_out/bench_mesh.json | cut -c1-1500
