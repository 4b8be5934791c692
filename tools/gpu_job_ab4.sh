#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
AB_OUT=gpurun_out/ab4_base.npy timeout 100 python tools/ab_flow.py > gpurun_out/ab4_base.json 2> gpurun_out/ab4_base.err; cat gpurun_out/ab4_base.json
SOFIMA_B200_LIB=$PWD/sofima_b200/_lib/ab/libV8.so AB_OUT=gpurun_out/ab4_v8.npy timeout 100 python tools/ab_flow.py > gpurun_out/ab4_v8.json 2> gpurun_out/ab4_v8.err; cat gpurun_out/ab4_v8.json
( SOFIMA_B200_LIB=$PWD/sofima_b200/_lib/ab/libV8.so timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_v8.log 2>&1; tail -2 gpurun_out/pytest_gpu_v8.log
