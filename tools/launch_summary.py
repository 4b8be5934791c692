"""Summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):

  python tools/launch_summary.py launches.csv "title" > profiles/launches_rN_summary.txt
"""
import collections, csv, io, re, sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
unit = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 's': 1e6}
agg = collections.OrderedDict()
for r in rows:
  if r['Metric Name'] != 'gpu__time_duration.sum':
    continue
  name = re.sub(r'\(.*$', '', r['Kernel Name']).replace('sofima::', '')[:62]
  us = float(r['Metric Value'].replace(',', '')) * unit[r['Metric Unit']]
  a = agg.setdefault(name, [0, 0.0])
  a[0] += 1
  a[1] += us
total = sum(a[1] for a in agg.values())
print(f'# ncu launch list summary ({sys.argv[2]})')
print('# per-launch device time, cold-cache and serialised: compare SHARES, not absolutes')
print(f'{"kernel":<62} {"launches":>8} {"total_us":>11} {"avg_us":>9} {"share":>7}')
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print(f'{name:<62} {n:>8} {us:>11.1f} {us / n:>9.2f} {100 * us / total:>6.1f}%')
