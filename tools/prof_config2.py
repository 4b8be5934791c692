import sys, os, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from sofima_b200 import stitch_elastic
tiles, cxm, cym = bench.synth_tile_grid(4, 4, bench.FLOW_TILE)
def strips():
  for axis, cm in ((0, cxm), (1, cym)):
    stitch_elastic.compute_flow_map(tiles, cm, axis, (160, 160), (40, 40), 1024)
strips()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); strips(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
