#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -2 gpurun_out/bench_n4.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_n4.json'))
print('flow',b['value'],'e2e',b['e2e']['value'],'mesh',b['mesh']['value'],b['mesh']['ms_per_step'],'mesh e2e',b['mesh']['e2e']['value'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/bench_n4_ref.json 2> gpurun_out/bench_n4_ref.err
cut -c1-300 gpurun_out/bench_n4_ref.json
