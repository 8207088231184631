import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.manual_seed(0)
n, g = 200000, 1000
A = torch.randn((n, g), device="cuda"); B = 0.4 * torch.randn((n, g), device="cuda") + 0.1 * A
ref = A.double().T @ B.double()
for chunk in (128, 256, 512, 1024, 2048):
    os.environ["SC_LEE_TC_CHUNK"] = str(chunk)
    L = eng.lee_gemm(A, B, g, impl=2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.lee_gemm(A, B, g, impl=2); e1.record(); torch.cuda.synchronize()
    err = (L.double() - ref).abs()
    print(f"chunk={chunk}: {e0.elapsed_time(e1):.2f} ms  max abs {err.max().item():.3e}  max rel-to-max {err.max().item()/ref.abs().max().item():.2e}  median abs {err.median().item():.2e}  (offdiag scale {ref.abs().median().item():.1f})", flush=True)
