"""Per-source-line stall samples of an .ncu-rep captured with --import-source on:
   python tools/ncu_lines.py rep.ncu-rep [top]
Aggregates the warp-stall samples of the SASS rows under each CUDA source line (all files of
the first kernel) and prints the top lines with their dominant stall reasons."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = '?', None
samples = collections.Counter()
stalls = collections.defaultdict(collections.Counter)
insts = collections.Counter()
text = {}
cur = None
seen_kernel = 0
for r in rows:
  if not r:
    continue
  if r[0] == 'Function Name':
    continue
  if r[0] == 'File Name':
    fname = r[1].split('/')[-1]
    continue
  if r[0] == 'Line No':
    hdr = r
    i_s, i_n = hdr.index('# Samples'), hdr.index('Instructions Executed')
    st_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    continue
  if hdr is None or len(r) < len(hdr):
    continue
  if r[0]:
    cur = (fname, int(r[0]))
    text[cur] = r[1].strip()
    continue
  try:
    n = int(r[i_s] or 0)
  except ValueError:
    continue
  samples[cur] += n
  try:
    insts[cur] += int(r[i_n] or 0)
  except ValueError:
    pass
  for i, name in st_cols:
    try:
      stalls[cur][name] += int(r[i] or 0)
    except ValueError:
      pass
tot = sum(samples.values()) or 1
print('total samples', tot, ' warp instructions', sum(insts.values()))
byfile = collections.Counter()
for (f, _), n in samples.items():
  byfile[f] += n
print('by file:', ', '.join(f'{f} {100*n/tot:.1f}%' for f, n in byfile.most_common()))
for key, n in samples.most_common(top):
  s = stalls[key]
  ss = sum(s.values()) or 1
  why = ', '.join(f'{k} {100*v/ss:.0f}%' for k, v in s.most_common(3))
  print(f'{100*n/tot:5.1f}%  {key[0]}:{key[1]:<4d} inst {insts[key]:>9d}  [{why}]  {text.get(key, "")[:70]}')
