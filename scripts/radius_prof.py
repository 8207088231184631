"""One C4-shaped radius graph build (for ncu launch lists / captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng, synthetic
n = 5_000_000
c = synthetic.coords_uniform(n, 1.2e5, 3)
cd = torch.from_numpy(c).cuda()
r = synthetic.radius_for_mean_degree(n, 1.2e5, 20.0)
for _ in range(2):
    g, _ = eng.radius_graph(cd, r)
torch.cuda.synchronize()
print("ok", g.nnz)
