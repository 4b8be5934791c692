"""DRAM traffic and device time of every kernel of ONE bench step, from an ncu metrics pass:

  python tools/step_traffic.py flow|mesh out.json      (on the GPU box; runs ncu itself)

flow: one 4096^2 tile pair (9801 patch pairs, 10 reference batches) through _FlowJob.run;
mesh: 30 FIRE steps on 2048^2 nodes.  The first (warm-up) pass of tools/prof_target.py is
skipped by counting launches.  ncu serialises kernels and flushes caches between passes, so the
times are cold-cache: use the SHARES and the bytes, not the absolute times."""
import collections, csv, io, json, os, subprocess, sys

which, out = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cmd = ['ncu', '--metrics', 'dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,'
       'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,'
       'smsp__issue_active.avg.pct_of_peak_sustained_active',
       '--clock-control', 'none', '--csv', sys.executable, os.path.join(root, 'tools', 'prof_target.py'), which]
res = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, PROF_MESH_ITERS='30'))
lines = res.stdout.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
launch = collections.OrderedDict()
for r in rows:
  launch.setdefault(r['ID'], {'name': r['Kernel Name']})[r['Metric Name']] = (
      float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
def to_bytes(v, u):
  return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
def to_us(v, u):
  return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1, 'nsecond': 1e-3, 's': 1e6}[u]
ours = [l for l in launch.values() if 'sofima' in l['name'] or any(
    k in l['name'] for k in ('cols_fast', 'rows_inv', 'rowspec', 'peak', 'rowcache_meta', 'pair_fused',
                             'fused_', 'mesh2d', 'pack2d', 'finalize2d', 'init_state', 'patch_'))]
# two identical passes were run (warm-up + measured): keep the second half
half = len(ours) // 2
step = ours[half:]
agg = collections.OrderedDict()
for l in step:
  name = l['name'].split('(')[0].replace('void ', '').replace('sofima::', '')
  a = agg.setdefault(name, {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0,
                            'time_us': 0.0, 'tensor_pipe_pct_max': 0.0, 'issue_active_pct_avg': 0.0})
  a['launches'] += 1
  a['dram_read_bytes'] += to_bytes(*l['dram__bytes_read.sum'])
  a['dram_write_bytes'] += to_bytes(*l['dram__bytes_write.sum'])
  a['time_us'] += to_us(*l['gpu__time_duration.sum'])
  a['tensor_pipe_pct_max'] = max(a['tensor_pipe_pct_max'],
                                 l['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0])
  a['issue_active_pct_avg'] += l['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]
for a in agg.values():
  a['issue_active_pct_avg'] /= a['launches']
tot = {'dram_bytes': sum(a['dram_read_bytes'] + a['dram_write_bytes'] for a in agg.values()),
       'time_us': sum(a['time_us'] for a in agg.values())}
json.dump({'what': f'{which}: one bench step under ncu (metrics pass, cold caches, serialised)',
           'kernels': agg, 'total': tot}, open(out, 'w'), indent=1)
print(json.dumps(tot))
for k, a in agg.items():
  print(f"{k[:60]:60s} n={a['launches']:3d} {a['time_us']:9.1f} us  rd {a['dram_read_bytes']/1e6:9.1f} MB  "
        f"wr {a['dram_write_bytes']/1e6:9.1f} MB  tensor {a['tensor_pipe_pct_max']:.1f}%  issue {a['issue_active_pct_avg']:.0f}%")
