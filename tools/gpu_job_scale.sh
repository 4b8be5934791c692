#!/bin/bash
# bench.py at N GPUs (both arms of interest) + BASELINE config 4 at the same N.
# usage: tools/gpu_job_scale.sh N
set -u
N=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
  python tools/config_runs.py config4 > gpurun_out/config4_r2_n1.json 2> gpurun_out/config4_r2_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tools/config_runs.py config4 > gpurun_out/config4_r2_n$N.json 2> gpurun_out/config4_r2_n$N.err
fi
tail -2 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r2_n$N.json'))
m = d['mesh']
print('N=$N flow', d['value'], 'e2e', d['e2e']['value'], '| mesh', m['value'], 'e2e', m['e2e']['value'],
      'ms/step', m['ms_per_step'], 'frac', m['roofline']['frac'], 'parity', m.get('parity', {}).get('bit_identical'))
try:
  c = [json.loads(l) for l in open("gpurun_out/config4_r2_n$N.json") if l.startswith("{")][-1]
  print('config4', c['patch_pairs_per_s'], c['seconds'], c['parity'])
except Exception as e:
  print('config4 failed', e)
PY
