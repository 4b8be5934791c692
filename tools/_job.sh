for v in "prev:SOFIMA_B200_LIB=tools/candidates/lib_prev.so" "new:X=1" "prev2:SOFIMA_B200_LIB=tools/candidates/lib_prev.so" "new2:X=1"; do
  name=${v%%:*}; kv=${v#*:}
  env $kv AB_OUT=gpurun_out/abt_$name.npy timeout 300 python tools/ab_flow.py > gpurun_out/abt_$name.json 2> gpurun_out/abt_$name.err
  echo "$name $(python -c "import json;d=json.load(open('gpurun_out/abt_$name.json'));print(d['ms_per_step'], d['kernel_ms_per_step'])")"
done
python - <<'PY'
import numpy as np
a, b = np.load('gpurun_out/abt_prev.npy'), np.load('gpurun_out/abt_new.npy')
print('outputs bit-identical:', np.array_equal(a, b, equal_nan=True))
PY
