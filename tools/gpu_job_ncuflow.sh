#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cols_fast|rows_inv_fast|rowspec_fast' -s 4 -c 4 -f -o gpurun_out/prof_flow2 python tools/prof_target.py flow > gpurun_out/ncu_flow2.log 2>&1
tail -3 gpurun_out/ncu_flow2.log
