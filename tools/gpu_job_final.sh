#!/bin/bash
# round-end style visit: parity tests, smoke, the default bench line, the reference arm,
# ncu launch list (never a bench number), warp path
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --mesh-iters 50 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
for m in lanczos linear nearest; do
  timeout 300 python tools/bench_warp.py --interpolation $m --out gpurun_out/warp_r2_$m.json > /dev/null 2> gpurun_out/warp_$m.err
done
cut -c1-300 gpurun_out/bench.json
