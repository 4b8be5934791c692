N=${1:-2}
mkdir -p gpurun_out/trace
SOFIMA_SHARD_TRACE=gpurun_out/trace/t$N timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi/run_sharded_mesh.py 2048 2048 1000 1 2>&1 | grep -v "^\*\*\*\|NCCL" | tail -3
python tools/shard_trace.py gpurun_out/trace/t$N --json gpurun_out/shard_trace_n$N.json | head -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/multi/run_sharded_mesh.py 2048 2048 1000 1 2>&1 | grep -v "^\*\*\*\|NCCL" | tail -1
