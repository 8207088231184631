"""Per-kernel roofline measurements of the non-headline rows (one GPU): z-score, lag + Moran
statistic (user order vs spatial order), value-permuting null (register-gather vs materialised),
spatial re-ordering.  Algorithmic bytes per SURVEY.md §8d.  Prints one JSON object.

    python scripts/bench_kernels.py [C4|C2] [sections]
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng, synthetic
import logging; logging.getLogger("spatialcore").setLevel(logging.ERROR)

PEAK = 6532.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

which = sys.argv[1] if len(sys.argv) > 1 else "C4"
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
def want(name): return only is None or name in only

CFG = {"C4": dict(n=5_000_000, g=1000, ext=1.2e5, graph="radius", deg=20.0, gen="uniform", seed=3),
       "C2": dict(n=500_000, g=400, ext=1e4, graph="knn", k=15, gen="mixture", seed=1),
       "C3": dict(n=200_000, g=1000, ext=6e3, graph="knn", k=6, gen="mixture", seed=2)}[which]


def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def roof(nbytes, ms):
    gbs = nbytes / (ms / 1e3) / 1e9
    return {"ms": round(ms, 3), "algorithmic_GB": round(nbytes / 1e9, 3), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 4)}


n, g = CFG["n"], CFG["g"]
out = {"config": which, "n": n, "g": g, "peak_GBps": PEAK}
c = synthetic.coords_uniform(n, CFG["ext"], CFG["seed"]) if CFG["gen"] == "uniform" else synthetic.coords_mixture(n, CFG["ext"], CFG["seed"])
cd = torch.from_numpy(c).cuda()
X = synthetic.expression_device(c, g, seed=CFG["seed"] * 1000, device="cuda")
if CFG["graph"] == "radius":
    r = synthetic.radius_for_mean_degree(n, CFG["ext"], CFG["deg"])
    graph, _ = eng.radius_graph(cd, r)
    out["graph_build_ms"] = timed(lambda: eng.radius_graph(cd, r))
else:
    graph, _, _ = eng.knn_graph(cd, CFG["k"])
    out["graph_build_ms"] = timed(lambda: eng.knn_graph(cd, CFG["k"]))
nnz = graph.nnz
out["nnz"] = nnz
ld = eng.padded_ld(g)

co = eng.spatial_order(cd)
out["spatial_order_ms"] = timed(lambda: eng.spatial_order(cd))
gs = eng.relabel_graph(graph, co)
out["graph_relabel_ms"] = timed(lambda: eng.relabel_graph(graph, co))

if want("zscore"):
    out["zscore_user_order"] = roof(12.0 * n * g, timed(lambda: eng.zscore_dense(X)))
    out["zscore_spatial_order"] = roof(12.0 * n * g, timed(lambda: eng.zscore_dense(X, rows=co.order)))

std_u = eng.zscore_dense(X)
if want("lag"):
    lag_bytes = 8.0 * n * g + 4.0 * nnz + 4.0 * n
    out["lag_user_order"] = roof(lag_bytes, timed(lambda: eng.lag_moran(graph, std_u.Z, g)))
    out["lag_stat_only_user_order"] = roof(4.0 * n * g + 4.0 * nnz + 4.0 * n, timed(lambda: eng.lag_moran(graph, std_u.Z, g, want_lag=False)))
num_u, den_u, _, _ = eng.lag_moran(graph, std_u.Z, g, want_lag=False)
del std_u
std = eng.zscore_dense(X, rows=co.order)
del X
if want("lag"):
    out["lag_spatial_order"] = roof(lag_bytes, timed(lambda: eng.lag_moran(gs, std.Z, g)))
    out["lag_stat_only_spatial_order"] = roof(4.0 * n * g + 4.0 * nnz + 4.0 * n, timed(lambda: eng.lag_moran(gs, std.Z, g, want_lag=False)))
    num_s, den_s, _, _ = eng.lag_moran(gs, std.Z, g, want_lag=False)
    out["lag_num_rel_diff_user_vs_spatial"] = float(((num_s - num_u).abs() / num_u.abs().clamp_min(1e-30)).max())

if want("lagsweep"):
    sweep = {}
    for q in (8, 16, 32):
        for chunk in (64, 128, 256, 512):
            os.environ["SC_LAG_Q"], os.environ["SC_LAG_CHUNK"] = str(q), str(chunk)
            sweep[f"q{q}_chunk{chunk}"] = [round(timed(lambda: eng.lag_moran(gs, std.Z, g)), 3),
                                           round(timed(lambda: eng.lag_moran(gs, std.Z, g, want_lag=False)), 3)]
    del os.environ["SC_LAG_Q"], os.environ["SC_LAG_CHUNK"]
    out["lag_sweep_ms_[with_lag,stat_only]"] = sweep

if want("lagtile"):
    # shared-memory tile lag (csrc/lag_tile.cu) against the L1-gather kernel
    res = {}
    gs.tiles = None
    num_ref, _, lag_ref, _ = eng.lag_moran(gs, std.Z, g)
    lag_bytes = 8.0 * n * g + 4.0 * nnz + 4.0 * n
    res["gather_lag"] = roof(lag_bytes, timed(lambda: eng.lag_moran(gs, std.Z, g)))
    res["gather_stat_only_ms"] = round(timed(lambda: eng.lag_moran(gs, std.Z, g, want_lag=False)), 3)
    res["tile_build_ms"] = round(timed(lambda: eng.tile_graph(gs)), 3)
    res["tile_lag"] = roof(lag_bytes, timed(lambda: eng.lag_moran(gs, std.Z, g)))
    res["tile_stat_only_ms"] = round(timed(lambda: eng.lag_moran(gs, std.Z, g, want_lag=False)), 3)
    num_g, _, lag_g, _ = eng.lag_moran(gs, std.Z, g)
    res["tile_lag_bitwise_equal_to_gather_kernel"] = bool(torch.equal(lag_g, lag_ref))
    res["tile_num_rel_diff"] = float(((num_g - num_ref).abs() / num_ref.abs().clamp_min(1e-30)).max())
    del lag_g, lag_ref
    ms = timed(lambda: eng.perm_null_values(gs, std.Z, g, 2, seed=1), reps=2) / 2
    res["tile_values_null"] = dict(ms_per_perm=round(ms, 3), gene_perms_per_s=round(g / (ms / 1e3), 1))
    out["lag_tiles"] = res

if want("values"):
    k1 = nnz / n + 1.0
    P = 4
    compulsory = 4.0 * n * (1.0 + k1 / g) * g * P  # HBM-compulsory bytes of P permutations x g genes
    ms = timed(lambda: eng.perm_null_values(gs, std.Z, g, P, seed=1), reps=2)
    out["values_null_materialised_spatial"] = dict(roof(compulsory, ms), gene_perms_per_s=round(g * P / (ms / 1e3), 1), ms_per_perm=round(ms / P, 3))
    os.environ["SC_PERM_VALUES_VARIANT"] = "gather"
    ms = timed(lambda: eng.perm_null_values(gs, std.Z, g, P, seed=1), reps=1)
    out["values_null_gather_spatial"] = dict(roof(compulsory, ms), gene_perms_per_s=round(g * P / (ms / 1e3), 1), ms_per_perm=round(ms / P, 3))
    del os.environ["SC_PERM_VALUES_VARIANT"]
    # per-cell counters (local Moran epilogue)
    _, _, _, loc = eng.lag_moran(gs, std.Z, g, want_lag=False, want_local=True)
    cnt = torch.zeros(std.Z.shape, dtype=torch.int32, device="cuda")
    ms = timed(lambda: eng.perm_null_values(gs, std.Z, g, P, seed=1, cell_obs=loc, cell_cnt=cnt), reps=2)
    out["values_null_materialised_cellcounts"] = dict(gene_perms_per_s=round(g * P / (ms / 1e3), 1), ms_per_perm=round(ms / P, 3))
    del loc, cnt

if want("rows"):
    _, _, lag, _ = eng.lag_moran(gs, std.Z, g)
    P = 64
    ms = timed(lambda: eng.perm_null_graph_rows(std.Z, lag, g, P, seed=1), reps=2)
    out["graph_rows_null"] = dict(roof(4.0 * n * (1 + 1.0 / 999) * g * P, ms), gene_perms_per_s=round(g * P / (ms / 1e3), 1))
print(json.dumps(out, indent=1))
