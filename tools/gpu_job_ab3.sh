#!/bin/bash
mkdir -p gpurun_out
AB_OUT=gpurun_out/ab3.npy timeout 200 python tools/ab_flow.py > gpurun_out/ab3.json 2> gpurun_out/ab3.err
cat gpurun_out/ab3.json; tail -2 gpurun_out/ab3.err
python - <<'PY'
import numpy as np
ref=np.load('gpurun_out/ab2_unsorted.npy'); x=np.load('gpurun_out/ab3.npy')
print('xy equal:', np.array_equal(ref[:,:2],x[:,:2],equal_nan=True), 'stats max rel', float(np.nanmax(np.abs(ref[:,2:]-x[:,2:])/np.maximum(np.abs(ref[:,2:]),1e-30))))
PY
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
