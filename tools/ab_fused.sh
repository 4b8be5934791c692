#!/bin/bash
# A/B of the flow path variants on one B200 (tools/ab_flow.py): three-kernel path vs the
# fused kernel with 3 / 4 thread groups; outputs compared bit for bit.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "three:SOFIMA_FLOW_FUSED=0" "fused3:SOFIMA_FLOW_FUSED=1 SOFIMA_FLOW_FUSED_GROUPS=3" "fused4:SOFIMA_FLOW_FUSED=1 SOFIMA_FLOW_FUSED_GROUPS=4"; do
  name=${v%%:*}; kv=${v#*:}
  env $kv AB_OUT=gpurun_out/abf_$name.npy timeout 300 python tools/ab_flow.py > gpurun_out/abf_$name.json 2> gpurun_out/abf_$name.err
  echo "$name rc=$? $(cat gpurun_out/abf_$name.json | cut -c1-600)"
  tail -3 gpurun_out/abf_$name.err
done
python - <<'PY'
import numpy as np
a = np.load('gpurun_out/abf_three.npy')
for n in ('fused3', 'fused4'):
  try:
    b = np.load(f'gpurun_out/abf_{n}.npy')
  except OSError as e:
    print(n, 'missing', e); continue
  same = np.array_equal(a, b, equal_nan=True)
  print(n, 'bit-identical to the three-kernel path:', same)
  if not same:
    d = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
    print('  first differences:', d[:8].tolist(), a[d[0][0]], b[d[0][0]])
PY
