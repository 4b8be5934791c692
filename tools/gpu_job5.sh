#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --path flow --no-cpu-baseline > gpurun_out/bench_flow.json 2> gpurun_out/bench_flow.err
tail -3 gpurun_out/bench_flow.err; cut -c1-1800 gpurun_out/bench_flow.json
