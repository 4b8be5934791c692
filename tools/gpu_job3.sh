#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/stitch_bench.py > gpurun_out/stitch_bench.json 2> gpurun_out/stitch_bench.err
tail -3 gpurun_out/stitch_bench.err; cut -c1-1500 gpurun_out/stitch_bench.json
