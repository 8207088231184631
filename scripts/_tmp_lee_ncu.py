import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.manual_seed(0)
n, g = 200000, 1000
A = torch.randn((n, g), device="cuda"); B = 0.4 * torch.randn((n, g), device="cuda") + 0.02 * A
for _ in range(3):
    eng.lee_gemm(A, B, g, impl=2); torch.cuda.synchronize()
