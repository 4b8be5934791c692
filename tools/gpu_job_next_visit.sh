#!/bin/bash
# First GPU visit of the next round: (1) the two pieces written without GPU time -- rigid
# tile-grid relaxation and masked 3-d correlation -- with full output (XPASS / XFAIL and the
# child's stderr), (2) the prepared kernel candidates of tools/candidates/ A/B'd against the
# in-tree library on the same box, parity suite on each candidate.  ~3 GPU-minutes.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tile_mesh_gpu.py tests/test_masked3d_gpu.py -m gpu -rxX -q \
  > gpurun_out/hedged_tests.log 2>&1; tail -15 gpurun_out/hedged_tests.log
L=$PWD/sofima_b200/_lib
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
AB_OUT=gpurun_out/cand_base.npy timeout 100 python tools/ab_flow.py > gpurun_out/cand_base.json 2> gpurun_out/cand_base.err
echo "base: $(cat gpurun_out/cand_base.json)"
for c in cols_product_in_smem_5_blocks rows_inv_reverse_pair_order; do
  rm -rf /tmp/cand && mkdir -p /tmp/cand/sofima_b200 /tmp/cand/include
  cp -r sofima_b200/csrc /tmp/cand/sofima_b200/ && cp include/sofima_b200.h /tmp/cand/include/
  (cd /tmp/cand && patch -p1 -s < $OLDPWD/tools/candidates/$c.patch) || { echo "$c: patch failed"; continue; }
  nvcc $F -c /tmp/cand/sofima_b200/csrc/flow.cu -o /tmp/cand/flow.o || { echo "$c: nvcc failed"; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /tmp/cand/lib$c.so $L/ctx.o $L/mesh.o $L/tile_mesh.o $L/warp.o /tmp/cand/flow.o
  SOFIMA_B200_LIB=/tmp/cand/lib$c.so AB_OUT=gpurun_out/cand_$c.npy timeout 100 python tools/ab_flow.py > gpurun_out/cand_$c.json 2> gpurun_out/cand_$c.err
  echo "$c: $(cat gpurun_out/cand_$c.json)"
  ( SOFIMA_B200_LIB=/tmp/cand/lib$c.so timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/cand_$c.pytest.log 2>&1
  tail -2 gpurun_out/cand_$c.pytest.log
done
python - <<'PY'
import numpy as np, glob
base = np.load('gpurun_out/cand_base.npy')
for f in sorted(glob.glob('gpurun_out/cand_*.npy')):
  x = np.load(f)
  print(f, 'identical to base:', np.array_equal(base, x, equal_nan=True))
PY
