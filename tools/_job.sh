timeout 900 python tools/config_runs.py config5 --depth 128 --render > gpurun_out/config5_r2_render.json 2> gpurun_out/config5_r2_render.err
tail -c 1500 gpurun_out/config5_r2_render.json; tail -5 gpurun_out/config5_r2_render.err
