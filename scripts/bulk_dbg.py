import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from spatialcore_b200 import engine as eng
g = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for n in [int(x) for x in sys.argv[1].split(',')]:
    A = torch.randn((n, eng.padded_ld(g)), device='cuda'); B = torch.randn_like(A)
    res = {}
    for v in ('8d', 'bulk8', 'bulk16'):
        os.environ['SC_PERM_ROWS_VARIANT'] = v
        try:
            s = eng.perm_null_graph_rows(A, B, g, 40, seed=3, perm_offset=5)
            torch.cuda.synchronize()
            res[v] = s
            print(n, v, 'ok', float(s.abs().max()), flush=True)
        except Exception as e:
            print(n, v, 'FAILED', str(e)[:200], flush=True); raise
    print(n, 'max diff bulk8', float((res['8d']-res['bulk8']).abs().max()), 'bulk16', float((res['8d']-res['bulk16']).abs().max()), flush=True)
    del A, B
