timeout 900 python -m pytest tests/test_mesh_gpu.py tests/test_stitch_gpu.py tests/test_relax_mesh_proc_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_parity_gpu.py tests/test_tile_mesh_gpu.py -x -q 2>&1 | tail -3
timeout 200 python /dev/stdin <<'PY' 2>&1 | cut -c1-140
import sys, os, json
sys.path.insert(0, os.getcwd() + '/tools'); sys.path.insert(0, os.getcwd())
import dev_mesh_bench as d
d.run(2048, 1, 1000, True, True)
d.run(102, 16, 1000, True, True)
PY
timeout 300 python tools/stitch_bench.py 2>&1 | tail -4 | cut -c1-300
