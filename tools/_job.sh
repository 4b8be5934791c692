timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2_n1.json'))
m = d['mesh']
print('flow', d['value'], 'e2e', d['e2e']['value'], d['roofline']['kernel_ms_per_step'])
print('mesh', m['value'], 'e2e', m['e2e']['value'], 'ms/step', m['ms_per_step'], 'frac', m['roofline']['frac'], m['roofline']['kernel_us_per_launch'], m['roofline']['kernel_us_per_launch_serialised'])
print('stitching', m['stitching']['value'], m['stitching']['us_per_integration_step'], 'config2', d['config2_flow']['value'])
PY
timeout 600 python tools/config_runs.py config5 --depth 128 > gpurun_out/config5_pdl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/config5_pdl.json')); print({k:d[k] for k in ('flow_seconds','mesh_seconds','mesh_us_per_step','mesh_steps')})"
