"""Per-step timeline of the persistent sharded mesh kernel (diagnostic).

Run the solver with SOFIMA_SHARD_TRACE=<prefix>; every rank dumps, for the last chunk, eight
%globaltimer stamps per step: block 0 at the top of the step (0), after the flag wait and the
FIRE update (1), after its tiles (2), after the block sum and the ticket (3); the block that
arrived last: on entry (4), after adding the rank's partials (5), after the system fence (6),
after the flags went out (7).  This script prints the mean duration of every phase in us:

  python tools/shard_trace.py <prefix> [--json out.json]
"""
import glob
import json
import sys

import numpy as np


def main():
  prefix = sys.argv[1]
  out = {}
  files = sorted(glob.glob(prefix + '.rank*'))
  for f in files:
    t = np.fromfile(f, dtype=np.uint64).reshape(-1, 8).astype(np.int64)
    t = t[5:-1]  # steady state
    nxt = np.fromfile(f, dtype=np.uint64).reshape(-1, 8).astype(np.int64)[6:]
    us = lambda a: float(np.mean(a)) / 1e3
    out[f.rsplit('.', 1)[1]] = {
        'step_period_us': us(nxt[:, 0] - t[:, 0]),
        'b0_wait_flags_and_fire_us': us(t[:, 1] - t[:, 0]),
        'b0_tiles_us': us(t[:, 2] - t[:, 1]),
        'b0_block_sum_ticket_us': us(t[:, 3] - t[:, 2]),
        'last_block_arrives_after_b0_us': us(t[:, 4] - t[:, 3]),
        'last_reduce_partials_us': us(t[:, 5] - t[:, 4]),
        'last_system_fence_us': us(t[:, 6] - t[:, 5]),
        'last_publish_us': us(t[:, 7] - t[:, 6]),
        'publish_to_next_step_start_us': us(nxt[:, 1] - t[:, 7]),
    }
  print(json.dumps(out, indent=1))
  if '--json' in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)


if __name__ == '__main__':
  main()
