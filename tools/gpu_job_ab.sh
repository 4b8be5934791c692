#!/bin/bash
# kernel A/B visit: HEAD flow kernels (A), streamlined rows_inv (B = in-tree default),
# B + packed fp32x2 codelets (C); parity suite on B and C.
mkdir -p gpurun_out
L=$PWD/sofima_b200/_lib
for v in A C; do
  SOFIMA_B200_LIB=$L/ab/lib$v.so AB_OUT=gpurun_out/ab_$v.npy timeout 200 python tools/ab_flow.py > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  cat gpurun_out/ab_$v.json; tail -2 gpurun_out/ab_$v.err
done
AB_OUT=gpurun_out/ab_B.npy timeout 200 python tools/ab_flow.py > gpurun_out/ab_B.json 2> gpurun_out/ab_B.err
cat gpurun_out/ab_B.json; tail -2 gpurun_out/ab_B.err
python - <<'PY'
import numpy as np
a,b,c=[np.load(f'gpurun_out/ab_{v}.npy') for v in 'ABC']
for n,x in (('B',b),('C',c)):
  print(n,'xy equal to A:', np.array_equal(a[:,:2],x[:,:2],equal_nan=True), 'stats max rel', float(np.nanmax(np.abs(a[:,2:]-x[:,2:])/np.maximum(np.abs(a[:,2:]),1e-30))))
PY
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_B.log 2>&1; tail -3 gpurun_out/pytest_gpu_B.log
( SOFIMA_B200_LIB=$L/ab/libC.so timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_C.log 2>&1; tail -3 gpurun_out/pytest_gpu_C.log
