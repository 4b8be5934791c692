#!/bin/bash
# A/B visit 2: packed codelets + streamlined rows_inv (now the in-tree default), with /
# without the x-major processing order, U sub-batch sizes (L2 residency of the product
# spectra); parity suite; ncu --set full of the column and inverse-row kernels.
mkdir -p gpurun_out
run() {  # name, env...
  n=$1; shift
  env "$@" AB_OUT=gpurun_out/ab2_$n.npy timeout 200 python tools/ab_flow.py > gpurun_out/ab2_$n.json 2> gpurun_out/ab2_$n.err
  echo "$n: $(cat gpurun_out/ab2_$n.json)"; tail -1 gpurun_out/ab2_$n.err
}
run sorted X=1
run unsorted SOFIMA_FLOW_SORT=0
run s96 SOFIMA_FLOW_SCRATCH_MB=96
run s160 SOFIMA_FLOW_SCRATCH_MB=160
run s256 SOFIMA_FLOW_SCRATCH_MB=256
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/ab_A.npy') if __import__('os').path.exists('gpurun_out/ab_A.npy') else None
ref=np.load('gpurun_out/ab2_unsorted.npy')
for n in ('sorted','s96','s160','s256'):
  x=np.load(f'gpurun_out/ab2_{n}.npy')
  print(n,'xy equal:', np.array_equal(ref[:,:2],x[:,:2],equal_nan=True), 'stats max rel', float(np.nanmax(np.abs(ref[:,2:]-x[:,2:])/np.maximum(np.abs(ref[:,2:]),1e-30))))
PY
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cols_fast|rows_inv_fast' -s 4 -c 4 -f -o gpurun_out/prof_flow3 python tools/prof_target.py flow > gpurun_out/ncu_flow3.log 2>&1; tail -2 gpurun_out/ncu_flow3.log
