#!/bin/bash
# compute-sanitizer over smoke() (every product kernel family on small inputs):
# memcheck (done in an earlier visit), racecheck (shared-memory hazards), synccheck
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 45 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
