timeout 900 python -m pytest tests/test_flow_gpu.py tests/test_masked3d_gpu.py tests/test_pipeline_gpu.py tests/test_stitch_gpu.py tests/test_missing_flow_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/prof_flow3d.py 2>&1 | tail -1
timeout 600 python tools/config_runs.py config5 --depth 128 --render > gpurun_out/config5_r2_render.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/config5_r2_render.json')); print({k:d[k] for k in ('flow_seconds','patch_pairs_per_s','mesh_seconds','render_seconds','problem_seconds')})"
