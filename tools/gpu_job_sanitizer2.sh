#!/bin/bash
# memcheck over the flow parity tests that run the cached-row-spectra fast path
mkdir -p gpurun_out
timeout 55 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python -m pytest tests/test_flow_gpu.py -q -x -m gpu -k "config1 or shared_row_spectra or kat or index_tables" > gpurun_out/sanitizer_memcheck_flow.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_flow.log
