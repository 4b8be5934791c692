timeout 900 python -m pytest tests/test_warp_cv_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_warp.py --interpolation lanczos --cpu-sections 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['value'], d['e2e']['value'], d['cpu_baseline'])"
