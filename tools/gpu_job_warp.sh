#!/bin/bash
# Warp path on the GPU box: tests, throughput, one ncu capture of the Lanczos kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_warp_cv_gpu.py -x -q 2>&1 | tail -3
for m in lanczos linear nearest; do
  timeout 300 python tools/bench_warp.py --interpolation $m --out gpurun_out/warp_r2_$m.json 2> gpurun_out/warp_$m.err | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_kernel -s 2 -c 1 \
  -o gpurun_out/ncu_r2_warp -f python tools/bench_warp.py --sections 4 --steps 1 --warmup 1 --cpu-sections 0 > gpurun_out/ncu_warp.log 2>&1
tail -2 gpurun_out/ncu_warp.log
