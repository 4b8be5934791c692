#!/bin/bash
# compute-sanitizer over the kernels added in round 2 and over smoke() (every product kernel
# family on small inputs): memcheck, racecheck (shared-memory hazards), synccheck
mkdir -p gpurun_out
run() {  # name, timeout, tool, command...
  local name=$1 to=$2 tool=$3; shift 3
  timeout $to compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 "$@" > gpurun_out/sanitizer_r2_$name.log 2>&1
  echo "$name ($tool) rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_r2_$name.log | tail -3
}
run memcheck_warp_cv 300 memcheck python -m pytest tests/test_warp_cv_gpu.py -q -x -m gpu -k "reference_outputs or variants or render_tiles or processor"
run memcheck_flowfilt 200 memcheck python -m pytest tests/test_flowfilt_gpu.py -q -x -m gpu
run memcheck_flow_tc 300 memcheck python -m pytest tests/test_flow_gpu.py -q -x -m gpu -k "tensor_core or fused or config1"
run memcheck_smoke 200 memcheck python -c "import __graft_entry__ as g; g.smoke()"
run racecheck_smoke 300 racecheck python -c "import __graft_entry__ as g; g.smoke()"
run synccheck_smoke 200 synccheck python -c "import __graft_entry__ as g; g.smoke()"
