#!/bin/bash
# A/B of the shared-row-spectra builder: rowspec_fast (fp32 FFT, CUDA cores) vs
# rowspec_tc_kernel (TMA + tcgen05.mma kind::i8, exact integer DFT); tools/ab_flow.py workload.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "fft:SOFIMA_FLOW_ROWSPEC_TC=0" "tc:SOFIMA_FLOW_ROWSPEC_TC=1"; do
  name=${v%%:*}; kv=${v#*:}
  env $kv AB_OUT=gpurun_out/abr_$name.npy timeout 300 python tools/ab_flow.py > gpurun_out/abr_$name.json 2> gpurun_out/abr_$name.err
  echo "$name rc=$? $(cat gpurun_out/abr_$name.json | cut -c1-700)"
  tail -3 gpurun_out/abr_$name.err
done
python - <<'PY'
import numpy as np
a, b = np.load('gpurun_out/abr_fft.npy'), np.load('gpurun_out/abr_tc.npy')
print('integer flow vectors identical:', np.array_equal(a[:, :2], b[:, :2], equal_nan=True))
ok = np.isfinite(a[:, 2]) & np.isfinite(b[:, 2])
rel = np.abs(1 / a[ok, 2] - 1 / b[ok, 2])
print('max |1/sharpness| difference:', float(rel.max()), ' max rel ratio diff:',
      float(np.nanmax(np.abs(a[ok, 3] - b[ok, 3]) / np.maximum(np.abs(a[ok, 3]), 1e-6))))
PY
