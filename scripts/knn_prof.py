"""One kNN build per configuration (for ncu launch lists / captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng, synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "u15"
cfg = {"u15": (5_000_000, 1.2e5, 15, "uniform"), "u6": (5_000_000, 1.2e5, 6, "uniform"),
       "m30": (2_000_000, 2e4, 30, "mixture"), "m15": (500_000, 1e4, 15, "mixture")}[which]
n, ext, k, gen = cfg
c = synthetic.coords_mixture(n, ext, 4) if gen == "mixture" else synthetic.coords_uniform(n, ext, 4)
cd = torch.from_numpy(c).cuda()
for _ in range(2):
    g, _, _ = eng.knn_graph(cd, k)
torch.cuda.synchronize()
print("ok", g.indices.shape)
