for m in lanczos nearest; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_kernel -s 2 -c 1 \
  -o gpurun_out/ncu_r2_warp_$m -f python tools/bench_warp.py --sections 4 --steps 1 --warmup 1 --cpu-sections 0 --interpolation $m > gpurun_out/ncu_warp.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
