timeout 900 python -m pytest tests/test_flow_gpu.py tests/test_masked3d_gpu.py tests/test_pipeline_gpu.py tests/test_stitch_gpu.py -x -q -m gpu -k "3d or 3 or pipeline or liconn or flow_map" 2>&1 | tail -3
timeout 300 python tools/prof_flow3d.py 2>&1 | tail -1
SOFIMA_FLOW3D_FAST=0 timeout 300 python tools/prof_flow3d.py 2>&1 | tail -1 | cut -c1-120
