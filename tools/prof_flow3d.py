"""Per-kernel device time of the 3-d patch correlation (config 5 geometry): 256 pairs of 80^3
patches of two [128, 512, 512] volumes.  Diagnostic, not a bench number."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sofima_b200 import _native, flow_field

ctx = _native.Context.get(0)
rng = np.random.default_rng(0)
a = torch.from_numpy(rng.integers(0, 255, (128, 512, 512), dtype=np.uint8)).cuda()
b = torch.roll(a, (2, -3, 4), (0, 1, 2)).contiguous()
calc = flow_field.JAXMaskedXCorrWithStatsCalculator()
run = lambda: calc.flow_field(a, b, (80, 80, 80), (40, 40, 40), batch_size=64)
out = run()
torch.cuda.synchronize()
ctx.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
rep = ctx.timing_report()
ctx.set_timing(False)
pairs = int(np.prod(out.shape[1:]))
print(json.dumps(dict(pairs=pairs, ms=e0.elapsed_time(e1), pairs_per_s=pairs / e0.elapsed_time(e1) * 1e3,
                      kernels={k: v for k, v in rep.items()})))
