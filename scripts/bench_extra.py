"""Secondary measurements (one GPU): graph builds, neighbourhood composition, Lee's L contraction,
value-permuting null, lag kernel, local Moran API.  Prints one JSON object."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import AnnDataLite, engine as eng, spatial, synthetic
import logging; logging.getLogger("spatialcore").setLevel(logging.ERROR)

only = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else None
def want(name): return only is None or name in only

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = {}
if want("graphs"):
    for name, n, ext, k, gen in (("C2_knn15_500k", 500_000, 1e4, 15, "mixture"), ("C5_knn30_2M", 2_000_000, 2e4, 30, "mixture"),
                                 ("knn15_5M", 5_000_000, 1.2e5, 15, "uniform"), ("knn6_5M", 5_000_000, 1.2e5, 6, "uniform")):
        c = synthetic.coords_mixture(n, ext, 4) if gen == "mixture" else synthetic.coords_uniform(n, ext, 4)
        cd = torch.from_numpy(c).cuda()
        out[name + "_ms"] = timed(lambda: eng.knn_graph(cd, k))
        out[name + "_with_dist_ms"] = timed(lambda: eng.knn_graph(cd, k, want_dist=True))
    c = synthetic.coords_uniform(5_000_000, 1.2e5, 3); cd = torch.from_numpy(c).cuda()
    r = synthetic.radius_for_mean_degree(5_000_000, 1.2e5, 20.0)
    out["C4_radius_5M_deg20_ms"] = timed(lambda: eng.radius_graph(cd, r))
    g, _ = eng.radius_graph(cd, r); out["C4_radius_nnz"] = g.nnz
    del cd, g

if want("nbhd"):
    n = 2_000_000
    c = synthetic.coords_mixture(n, 2e4, 4); lab = synthetic.patchy_labels(c, 30, 5)
    cd = torch.from_numpy(c).cuda(); ld = torch.from_numpy(lab).cuda()
    def nb():
        _, _, prof = eng.knn_graph(cd, 30, labels=ld, n_types=30, want_idx=False)
        eng.profile_normalize(prof, True)
    out["C5_nbhd_knn30_2M_T30_fused_ms"] = timed(nb)
    a = AnnDataLite(np.zeros((n, 1), np.float32), obsm={"spatial": c})
    import pandas as pd
    a.obs = pd.DataFrame({"ct": pd.Categorical.from_codes(lab, [f"t{i:02d}" for i in range(30)])})
    t0 = time.perf_counter(); spatial.compute_neighborhood_profile(a, "ct", k=30); out["C5_nbhd_api_e2e_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); spatial.compute_neighborhood_profile(a, "ct", k=30); out["C5_nbhd_api_e2e_s_2nd"] = time.perf_counter() - t0
    del cd, ld

if want("niches"):
    # C5: 2 M cells x 30 types, k=30 profiles -> 8 niches (n_init=10, max_iter=300 like the reference default)
    n = 2_000_000
    c = synthetic.coords_mixture(n, 2e4, 4); lab = synthetic.patchy_labels(c, 30, 5)
    cd = torch.from_numpy(c).cuda(); ld = torch.from_numpy(lab).cuda()
    _, _, prof = eng.knn_graph(cd, 30, labels=ld, n_types=30, want_idx=False)
    eng.profile_normalize(prof, True)
    from spatialcore_b200.spatial import niches
    torch.cuda.synchronize(); t0 = time.perf_counter()
    labels, cent, inertia, n_iter = niches.kmeans_fit(prof, 8, 10, 300, 0)
    torch.cuda.synchronize(); out["C5_niches_kmeans_2M_x30_k8_ninit10_s"] = time.perf_counter() - t0
    out["C5_niches_inertia"] = inertia
    km = eng.KMeansDevice(prof, 8)
    ms = timed(lambda: km.assign(cent))
    out["C5_kmeans_assign_pass_ms"] = ms
    out["C5_kmeans_assign_pass_GBps"] = (4.0 * n * 30 + 8.0 * n) / (ms / 1e3) / 1e9
    from sklearn.cluster import KMeans
    ph = prof.cpu().numpy()
    sub = ph[:: 10]
    t0 = time.perf_counter(); km_cpu = KMeans(n_clusters=8, init="k-means++", n_init=10, max_iter=300, random_state=0).fit(sub)
    out["cpu_sklearn_kmeans_200k_sample_s"] = time.perf_counter() - t0
    out["cpu_sklearn_cores"] = os.cpu_count()
    del cd, ld, prof

if want("lee"):
    n, g = 200_000, 1000
    c = synthetic.coords_mixture(n, 6e3, 2); cd = torch.from_numpy(c).cuda()
    X = synthetic.expression_device(c, g, 2)
    graph, _, _ = eng.knn_graph(cd, 6)
    std = eng.zscore_dense(X); _, _, lag, _ = eng.lag_moran(graph, std.Z, g)
    for impl in (1, 2):
        try:
            ms = timed(lambda: eng.lee_gemm(std.Z, lag, g, impl=impl))
            out[f"C3_lee_gemm_impl{impl}_ms"] = ms
            out[f"C3_lee_gemm_impl{impl}_useful_tflops"] = 2.0 * n * g * g / (ms / 1e3) / 1e12
        except Exception as e:
            out[f"C3_lee_gemm_impl{impl}"] = "unavailable: " + str(e)[:80]
    # all-pairs permutation null at C3: per permutation one row gather + lag pass + contraction
    co = eng.spatial_order(cd); gs = eng.relabel_graph(graph, co); std_s = eng.zscore_dense(X, rows=co.order)
    def one_perm():
        Zp = eng.gather_rows(std_s.Z, eng.philox_permutation(1, 0, n))
        _, _, lp, _ = eng.lag_moran(gs, Zp, g)
        eng.lee_gemm(std_s.Z, lp, g, impl=2)
    out["C3_lee_all_pairs_ms_per_permutation"] = timed(one_perm)
    L = eng.lee_gemm(std.Z, lag, g, impl=1)
    ref = (std.Z[:, :g].double().T @ lag[:, :g].double())
    out["C3_lee_impl1_max_abs_err_vs_fp64_torch"] = float((L.double() - ref).abs().max())
    del X, std, lag

if want("values"):
    for name, n, g, k, ext in (("C2", 500_000, 400, 15, 1e4), ("C1x", 100_000, 50, 6, 3e3)):
        c = synthetic.coords_mixture(n, ext, 1); cd = torch.from_numpy(c).cuda()
        X = synthetic.expression_device(c, g, 1)
        graph, _, _ = eng.knn_graph(cd, k)
        std = eng.zscore_dense(X)
        out[f"{name}_lag_moran_ms"] = timed(lambda: eng.lag_moran(graph, std.Z, g))
        P = 8
        ms = timed(lambda: eng.perm_null_values(graph, std.Z, g, P, seed=1), reps=2)
        out[f"{name}_values_null_gene_perms_per_s"] = g * P / (ms / 1e3)
        out[f"{name}_values_null_ms_per_perm"] = ms / P
        ms = timed(lambda: eng.perm_null_graph_rows(std.Z, eng.lag_moran(graph, std.Z, g)[2], g, 64, seed=1), reps=2)
        out[f"{name}_graph_rows_null_gene_perms_per_s_incl_lag"] = g * 64 / (ms / 1e3)
        del X, std

if want("local"):
    n, g = 200_000, 20
    c = synthetic.coords_mixture(n, 6e3, 1)
    X = synthetic.expression_device(c, g, 1).cpu().numpy()
    a = AnnDataLite(X, obsm={"spatial": c})
    for src in ("philox", "replay"):
        t0 = time.perf_counter(); spatial.local_morans_i(a, n_permutations=99, perm_source=src); out[f"local_morans_200k_x20_P99_{src}_s"] = time.perf_counter() - t0
if want("distances"):
    # calculate_domain_distances kernels: 1-NN of 1 M source cells among 1 M target cells; min/mean over 100 k x 100 k pairs
    rng = np.random.default_rng(2)
    T = rng.uniform(0, 2e4, (1_000_000, 2)); Q = rng.uniform(0, 2e4, (1_000_000, 2))
    Td, Qd = torch.from_numpy(T).cuda(), torch.from_numpy(Q).cuda()
    out["cross_nn_1Mx1M_ms"] = timed(lambda: eng.cross_nn(Td, Qd))
    from scipy.spatial import cKDTree
    t0 = time.perf_counter(); cKDTree(T).query(Q, k=1); out["cpu_ckdtree_1nn_1Mx1M_1thread_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); cKDTree(T).query(Q, k=1, workers=-1); out["cpu_ckdtree_1nn_1Mx1M_allcores_s"] = time.perf_counter() - t0
    A, B = Td[:100_000].contiguous(), Qd[:100_000].contiguous()
    out["pairwise_min_mean_100kx100k_ms"] = timed(lambda: eng.pairwise_reduce(A, B))
    from scipy.spatial.distance import cdist
    t0 = time.perf_counter(); D = cdist(T[:10_000], Q[:10_000]); D.min(); D.mean(); out["cpu_cdist_10kx10k_s"] = time.perf_counter() - t0

if want("local_ref"):
    # the one timing the reference publishes (docs/spatial/spatial_stats.md:202-215): local_morans_i on the
    # 366 938-cell CosMx colon vignette, 5 / 10 / 20 genes, 10 permutations: ~69 / 52 / 80 s (batched call)
    n = 366_938
    c = synthetic.coords_mixture(n, 8e3, 9)
    X = synthetic.expression_device(c, 20, 9).cpu().numpy()
    for g in (5, 10, 20):
        a = AnnDataLite(X[:, :g].copy(), obsm={"spatial": c})
        spatial.local_morans_i(a, n_permutations=10)  # warm
        t0 = time.perf_counter(); spatial.local_morans_i(a, n_permutations=10); out[f"local_morans_366938_x{g}_P10_s"] = time.perf_counter() - t0

print(json.dumps(out, indent=1))
