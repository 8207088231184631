import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.manual_seed(0)
for (n, g) in [(64, 32), (1000, 50), (5000, 300), (20000, 1000), (200000, 1000)]:
    ld = eng.padded_ld(g)
    A = torch.zeros((n, ld), device="cuda"); B = torch.zeros((n, ld), device="cuda")
    A[:, :g] = torch.randn((n, g), device="cuda"); B[:, :g] = 0.4 * torch.randn((n, g), device="cuda") + 0.1 * A[:, :g]
    ref = A[:, :g].double().T @ B[:, :g].double()
    L1 = eng.lee_gemm(A, B, g, impl=1); torch.cuda.synchronize()
    L2 = eng.lee_gemm(A, B, g, impl=2); torch.cuda.synchronize()
    e1 = (L1.double() - ref).abs().max().item(); e2 = (L2.double() - ref).abs().max().item()
    rel2 = ((L2.double() - ref).abs() / ref.abs().clamp_min(1e-30)).median().item()
    print(f"n={n} g={g}: impl1 max abs err {e1:.3e}  impl2 max abs err {e2:.3e} (scale sqrt(n)={n**0.5:.0f}, max|L|={ref.abs().max().item():.1f}, median rel {rel2:.2e})", flush=True)
    if n >= 20000:
        for impl in (1, 2):
            e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.lee_gemm(A, B, g, impl=impl); torch.cuda.synchronize()
            e0.record(); eng.lee_gemm(A, B, g, impl=impl); e1_.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1_)
            print(f"   impl{impl}: {ms:.2f} ms  {2.0*n*g*g/(ms/1e3)/1e12:.1f} useful TFLOP/s", flush=True)
