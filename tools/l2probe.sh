M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for args in "0 420 0" "1 420 0" "0 420 400" "1 420 400" "0 210 400" "1 210 400" "0 420 100" "1 420 100"; do
  echo "== $args"
  ncu $M --clock-control none -s 1 -c 1 tests/gpu_micro/_build/l2_rewrite_probe $args 2>&1 | grep -E "variant|dram__|duration|hit_rate"
done
