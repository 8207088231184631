"""Accuracy / time of the tensor-core Lee contraction (impl=2) against the TMEM accumulation length
(SC_LEE_TC_CHUNK cells per accumulator) on uncorrelated data (the hard case: entries are cancellation-dominated).
Error is reported in units of u * sum_i |a_i b_i| (u = 2^-24), the backward-error scale of an FP32 evaluation."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.manual_seed(0)
n, g = 200000, 1000
A = torch.randn((n, g), device="cuda"); B = 0.4 * torch.randn((n, g), device="cuda") + 0.02 * A
ref = A.double().T @ B.double()
mag = A.double().abs().T @ B.double().abs()
u = 2.0 ** -24
for chunk in (32, 64, 128, 256, 512):
    os.environ["SC_LEE_TC_CHUNK"] = str(chunk)
    L = eng.lee_gemm(A, B, g, impl=2); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.lee_gemm(A, B, g, impl=2); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    err = (L.double() - ref).abs()
    cond = (err / (u * mag)).flatten()
    rel = (err / ref.abs().clamp_min(1e-300)).flatten()
    q = torch.tensor([0.5, 0.99], device="cuda", dtype=torch.float64)
    idx = torch.randint(0, cond.numel(), (200000,), device="cuda")
    cq, rq = torch.quantile(cond[idx], q), torch.quantile(rel[idx], q)
    print(f"chunk={chunk}: {sorted(ts)[2]:.2f} ms  err/(u*sum|terms|) p50 {cq[0]:.3f} p99 {cq[1]:.3f} max {cond.max():.2f}  rel p50 {rq[0]:.2e} p99 {rq[1]:.2e}  above 1e-5: {(rel > 1e-5).double().mean():.3f}", flush=True)
