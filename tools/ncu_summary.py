"""Summarises an .ncu-rep (raw page + SASS page) into text: python tools/ncu_summary.py rep [kernel-index]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
  d = dict(zip(hdr, r))
  print('==', d.get('Kernel Name', '')[:90])
  for w in want:
    if w in d: print(f'   {w:62s} {d[w]:>16s} {units[hdr.index(w)]}')
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
names = [rows[i - 1][1] if i > 0 and len(rows[i-1]) > 1 else '?' for i in starts]
for bi, st in enumerate(starts):
  en = starts[bi + 1] - 1 if bi + 1 < len(starts) else len(rows)
  hdr = rows[st]
  ia, ie = hdr.index('Source'), hdr.index('Instructions Executed')
  stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
  mix, stalls, tot = collections.Counter(), collections.Counter(), 0
  for r in rows[st + 1:en]:
    if len(r) < len(hdr): continue
    try: n = int(r[ie])
    except ValueError: continue
    toks = r[ia].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    mix[op] += n; tot += n
    for i in stall_cols:
      try: stalls[hdr[i]] += int(r[i])
      except ValueError: pass
  print('== SASS', names[bi][:80], 'warp-insts', tot)
  print('   mix:', ', '.join(f'{k} {100*v/tot:.1f}%' for k, v in mix.most_common(14)))
  ssum = sum(stalls.values()) or 1
  print('   stalls:', ', '.join(f'{k[6:]} {100*v/ssum:.1f}%' for k, v in stalls.most_common(8)))

def by_line(rep, top=40):
  """Instruction counts aggregated per CUDA source line (needs -lineinfo)."""
  out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr = None
  agg, cur, tot = collections.Counter(), None, 0
  src = {}
  for r in rows:
    if r and r[0] == 'Line No':
      hdr = r
      ie = hdr.index('Instructions Executed')
      if any(agg.values()): break   # first kernel only
      continue
    if hdr is None or len(r) <= ie: continue
    if r[0]:
      cur = r[0]; src[cur] = r[1]
    else:
      try: n = int(r[ie])
      except ValueError: continue
      agg[cur] += n; tot += n
  print('== per-line warp instructions (first kernel), total', tot)
  for line, n in agg.most_common(top):
    print(f'   {100*n/tot:5.1f}%  L{line:>4s}  {src[line].strip()[:100]}')

if len(sys.argv) > 2 and sys.argv[2] == 'lines':
  by_line(rep)
