timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/config_runs.py config5 --depth 128 --render --problems 32 > gpurun_out/config5_r2_n8.json 2> gpurun_out/config5_r2_n8.err
grep '^{' gpurun_out/config5_r2_n8.json | tail -1 | cut -c1-2500; tail -3 gpurun_out/config5_r2_n8.err
