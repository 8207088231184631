import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.set_printoptions(precision=4, linewidth=200)
n, g = 64, 40
ld = eng.padded_ld(g)
A = torch.zeros((n, ld), device="cuda"); B = torch.zeros((n, ld), device="cuda")
# structured: A[i,x] = 1 if i == x (identity-like), B[i,y] = y + 100*i  -> L[x,y] = B[x,y]
for i in range(min(n, g)): A[i, i] = 1.0
B[:, :g] = torch.arange(g, device="cuda")[None, :].float() + 100.0 * torch.arange(n, device="cuda")[:, None].float()
ref = A[:, :g].double().T @ B[:, :g].double()
L2 = eng.lee_gemm(A, B, g, impl=2); torch.cuda.synchronize()
print("ref[:6,:8]\n", ref[:6, :8]); print("tc[:6,:8]\n", L2[:6, :8])
print("nonzero count", int((L2 != 0).sum()), "of", L2.numel(), " max", float(L2.abs().max()))
nz = (L2 != 0).nonzero()
print("first nonzeros:", nz[:10].tolist())
print("row 33..36 cols 0..8\n", L2[33:37, :8], "\nref\n", ref[33:37, :8])
