import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
torch.manual_seed(0)
n, g = 200000, 1000
A = torch.randn((n, g), device="cuda"); B = 0.4 * torch.randn((n, g), device="cuda") + 0.02 * A
def run(a=A, b=B, gg=g, reps=7):
    L = eng.lee_gemm(a, b, gg, impl=2); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.lee_gemm(a, b, gg, impl=2); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return L, sorted(ts)[reps // 2]
# small first (a hang here costs little)
a = torch.randn((4096, 256), device="cuda"); b = torch.randn((4096, 256), device="cuda")
os.environ["SC_LEE_TC_CTA2"] = "0"; L1, _ = run(a, b, 256, 1)
os.environ["SC_LEE_TC_CTA2"] = "1"; L2, _ = run(a, b, 256, 1)
print("small 4096x256: pair == single:", bool(torch.equal(L1, L2)), float((L1 - L2).abs().max()), flush=True)
os.environ["SC_LEE_TC_CTA2"] = "0"; L1, t1 = run()
print("single CTA", round(t1, 3), "ms", flush=True)
os.environ["SC_LEE_TC_CTA2"] = "1"; L2, t2 = run()
print("CTA pair  ", round(t2, 3), "ms; bit-identical:", bool(torch.equal(L1, L2)), float((L1 - L2).abs().max()), flush=True)
r64 = A.double().T @ B.double()
rel = ((L2.double() - r64).abs() / r64.abs().clamp_min(1e-300)).flatten()
print("rel err p50 %.2e" % float(torch.quantile(rel[torch.randint(0, rel.numel(), (200000,), device="cuda")], 0.5)))
for (nn, gg) in ((20011, 333), (70001, 1030), (513, 40), (8, 5)):
    a = torch.randn((nn, eng.padded_ld(gg)), device="cuda"); b = torch.randn((nn, eng.padded_ld(gg)), device="cuda")
    a[:, gg:] = 0; b[:, gg:] = 0
    L3 = eng.lee_gemm(a, b, gg, impl=2); R3 = (a.double().T @ b.double())[:gg, :gg]
    print((nn, gg), "max abs err", float((L3.double()[:gg, :gg] - R3).abs().max()), "scale", float(R3.abs().max()), flush=True)
