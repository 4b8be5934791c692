timeout 300 python -m pytest tests/test_warp_cv_gpu.py -x -q 2>&1 | tail -3
for m in lanczos linear nearest; do
  timeout 300 python tools/bench_warp.py --interpolation $m --cpu-sections 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][-8:], d['kernel_ms'], d['value'], d['e2e']['value'], d['cpu_baseline'])"
done
