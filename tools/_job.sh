timeout 600 python -m pytest tests/test_flow_gpu.py -x -q -m gpu -k "3d" 2>&1 | tail -8
