"""Bandwidth of the graph-row null kernel vs row width (genes per rank), one GPU."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
P = 64
for g in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "125,250,400,500,1000").split(",")]:
    ld = eng.padded_ld(g)
    A = torch.randn((n, ld), device="cuda"); B = torch.randn_like(A)
    for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["bulk16", "bulk8"]):
        os.environ["SC_PERM_ROWS_VARIANT"] = v
        eng.perm_null_graph_rows(A, B, g, 16, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.perm_null_graph_rows(A, B, g, P, seed=1); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gbs = 4.0 * n * g * P / (ms / 1e3) / 1e9
        print(f"n={n} g={g} ld={ld} {v}: {ms:.1f} ms  {gbs:.0f} GB/s algorithmic ({gbs/6532.2:.3f} of peak)", flush=True)
    del A, B
