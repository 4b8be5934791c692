timeout 900 python -m pytest tests/test_relax_mesh_proc_gpu.py tests/test_plugins.py tests/test_stitch_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -3
