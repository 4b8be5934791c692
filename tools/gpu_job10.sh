#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | tail -3
tail -2 gpurun_out/bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.json'))
print(b['value'], b.get('config2_flow'), b['mesh']['value'])
PY
