import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from sofima_b200 import flow_field as ff
rng = np.random.default_rng(0)
img = rng.integers(0, 255, (512, 512), dtype=np.uint8)
pre, post = img[:400, :400].copy(), img[3:403, 5:405].copy()
out = ff.JAXMaskedXCorrWithStatsCalculator().flow_field(pre, post, 160, 40, batch_size=16)
torch.cuda.synchronize()
print(out[:2, :3, :3])
