#!/bin/bash
# 2-GPU check of the sharded mesh (bit-exact vs single GPU) + 2-GPU bench line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi/run_sharded_mesh.py 512 384 200 > gpurun_out/sharded2.json 2> gpurun_out/sharded2.err
tail -2 gpurun_out/sharded2.err; cat gpurun_out/sharded2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi/run_sharded_mesh.py 2048 2048 300 > gpurun_out/sharded2_big.json 2>> gpurun_out/sharded2.err
cat gpurun_out/sharded2_big.json | cut -c1-600
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -2 gpurun_out/bench_n2.err; cut -c1-2500 gpurun_out/bench_n2.json
