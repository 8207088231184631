#!/usr/bin/env python
"""Headline benchmark: gene-permutations/s of global Moran's I (graph-row permutation null, the
null behind ``morans_i``) including the neighbour-graph build, on synthetic CosMx/Xenium-shaped data.

    python bench.py --gpus 1 --steps 3 --warmup 3                 # our arm (default workload C4)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1  # reference CPU arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N ...

One "step" = one pass of the whole hot path over the workload: graph build -> z-score -> lag +
Moran statistic -> graph moments -> P permutations -> per-gene p-values.  ``value`` times it with
inputs resident in HBM (CUDA events, max over ranks); ``e2e`` times the public API
``spatialcore_b200.spatial.morans_i`` on HOST arrays (pinned), H2D/D2H inside the timed region.
Multi-GPU (strong scaling of the fixed workload): ranks form gene blocks (>= 500 genes each) x
permutation groups; one all-reduce of the [4, G] null summary inside a gene block and one all-gather
of per-gene results at the end.  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: cells, genes, graph, permutations, extent, coordinate generator
    "C4": dict(n=5_000_000, g=1000, graph="radius", degree=20.0, k=None, perms=999, extent=120_000.0, coords="uniform",
               desc="CosMx-scale 5M cells x 1000 genes, radius graph (mean degree ~20), Moran's I, 999 permutations"),
    "C2": dict(n=500_000, g=400, graph="knn", degree=None, k=15, perms=999, extent=10_000.0, coords="mixture",
               desc="Xenium-scale 500k cells x 400 genes, kNN k=15, Moran's I, 999 permutations"),
    "C1": dict(n=10_000, g=50, graph="knn", degree=None, k=6, perms=99, extent=1_000.0, coords="uniform",
               desc="10k cells x 50 genes, kNN k=6, Moran's I, 99 permutations"),
}
METRIC = "gene-perms/sec Moran's I @5M cells x 1k genes (graph build + statistic + 999-permutation null)"  # BASELINE.json's metric, quoted on C4
UNIT = "gene-perms/s"


def metric_name(workload: str) -> str:
    """BASELINE.json's metric; the other workloads name their own size."""
    if workload == "C4":
        return METRIC
    w = WORKLOADS[workload]
    return f"gene-perms/sec Moran's I @{w['n']} cells x {w['g']} genes (graph build + statistic + {w['perms']}-permutation null)"


def workload_config(name: str, w: dict, radius) -> dict:
    """The ``config`` object of the JSON line -- identical for the b200 and the reference arm."""
    return {"workload": f"{name}: {w['desc']}", "n_cells": w["n"], "n_genes": w["g"], "n_permutations": w["perms"],
            "graph": w["graph"], "radius": radius, "k": w["k"], "null": "graph_rows (squidpy semantics)",
            "l2": "inputs (Z, lag: 4*N*G bytes each) far larger than the 126 MB L2; no flush needed"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SC_BENCH_WORKLOAD", "C4"), choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-values", action="store_true", help="skip the value-permuting-null leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the Lee's L / neighbourhood / kNN legs (BASELINE configs 2, 3, 5)")
    ap.add_argument("--seed", type=int, default=3)
    return ap.parse_args()


def make_coords(w, seed):
    from spatialcore_b200 import synthetic

    if w["coords"] == "uniform":
        return synthetic.coords_uniform(w["n"], w["extent"], seed)
    return synthetic.coords_mixture(w["n"], w["extent"], seed)


def workload_radius(w):
    from spatialcore_b200 import synthetic

    return synthetic.radius_for_mean_degree(w["n"], w["extent"], w["degree"]) if w["graph"] == "radius" else None


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time() + 1.0
        while not self.samples and time.time() < t_end:  # a leg shorter than nvidia-smi's start-up: wait for one sample
            time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------


def device_step(engine, ac, coords_dev, X_dev, w, radius, seed, events=None, perm_range=None, group=None):
    """One pass of the hot path with inputs resident in HBM.  Returns (I, p_value) device tensors."""
    import torch

    def mark(name):
        if events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.append((name, ev))

    n, g = X_dev.shape
    P = w["perms"]
    mark("start")
    if radius is not None:
        graph, _ = engine.radius_graph(coords_dev, radius, device=coords_dev.device)
    else:
        graph, _, _ = engine.knn_graph(coords_dev, w["k"], device=coords_dev.device)
    mark("graph")
    co = engine.spatial_order(coords_dev, device=coords_dev.device)
    graph_s = engine.relabel_graph(graph, co, tiles=False)
    mark("reorder")
    if engine.tile_rows() > 0:
        engine.tile_graph(graph_s)  # neighbour unions + word lists of the shared-memory lag kernel
    mark("tiles")
    std = engine.zscore_dense(X_dev, rows=co.order)
    mark("zscore")
    num, den, lag, _ = engine.lag_moran(graph_s, std.Z, g, want_lag=True)
    mark("lag")
    s0, s1, s2 = engine.graph_moments(graph_s)  # label-invariant; reverse-edge lookups are local in spatial order
    mark("moments")
    scale = (float(n) / s0) / den
    I = num * scale
    null = ac.MoranNull(g, X_dev.device)
    ac.moran_graph_rows_null(std.Z, lag, g, scale, I, P, seed, "philox", null, perm_range or (0, P))
    mark("perms")
    if group is not None:
        from spatialcore_b200 import distributed as dist_util

        dist_util.all_reduce_null(null, group)  # [4, G_block] FP64 over the ranks sharing this gene block
    c = torch.minimum(null.cnt_ge, P - null.cnt_ge)
    p_value = (c + 1).double() / (P + 1)
    return I, p_value


def _timed_ms(torch, fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def values_null_leg(spatial, AnnDataLite, engine, coords, coords_dev, X_dev, w, radius, seed, gpu_index):
    """Secondary headline (the metric is reported for both nulls): gene-perms/s of the
    VALUE-permuting null (the reference's own ``local_morans_i`` / Lee's L scheme) on the same workload,
    through the public API -- ``morans_i(null_mode="values")`` on a device-resident expression matrix, Philox
    permutations.  Two calls (P and 3P permutations) separate the per-permutation cost from the prologue."""
    import torch

    n, g = X_dev.shape
    names = [f"g{i}" for i in range(g)]
    sampler = ClockSampler(gpu_index)

    def call(P):
        adata = AnnDataLite(X_dev, obsm={"spatial": coords}, var_names=names)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        spatial.morans_i(adata, n_neighbors=w["k"] or 6, n_permutations=P, seed=seed, radius=radius, perm_source="philox",
                         null_mode="values", write_graph=False, shard="none", device=X_dev.device)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    call(2)  # warm-up
    sampler.start()
    P1, P3 = 8, 24
    t1, t3 = call(P1), call(P3)
    clocks = sampler.stop()
    per_perm = (t3 - t1) / (P3 - P1)
    if radius is not None:
        nnz = int(engine.radius_graph(coords_dev, radius, device=coords_dev.device)[0].nnz)
    else:
        nnz = n * w["k"]
    k1 = nnz / n + 1.0
    return {"value": round(g / per_perm, 1), "unit": UNIT, "ms_per_permutation": round(per_perm * 1e3, 3),
            "api": "spatialcore_b200.spatial.morans_i(null_mode='values', perm_source='philox') on a device-resident X",
            "api_seconds": {f"P={P1}": round(t1, 3), f"P={P3}": round(t3, 3)},
            "value_whole_call_P24": round(g * P3 / t3, 1),
            "null": "values (reference's own scheme: permute z, re-apply W)",
            "kernel": "lag_tile_kernel (shared-memory tiles, permutation applied while staging)",
            "bound": "shared-memory crossbar (LDS), SURVEY.md §8d: not HBM",
            "hbm_compulsory_gbs": round(4.0 * n * (1.0 + k1 / g) * g / per_perm / 1e9, 1),
            "lds_gather_gbs": round(4.0 * nnz * g / per_perm / 1e9, 1), "clocks": clocks}


def lee_leg(spatial, AnnDataLite, engine, synthetic, dev, gpu_index):
    """BASELINE config 3: Lee's L over all gene pairs, 200k cells x 1k genes, kNN k=6 -- the dense
    Z^T (W Z) contraction on tensor cores (tcgen05 kind::tf32, 3xTF32)."""
    import torch

    n, g, k = 200_000, 1000, 6
    coords = synthetic.coords_mixture(n, 6e3, 2)
    cd = torch.from_numpy(coords).to(dev)
    X = synthetic.expression_device(coords, g, 2, device=dev)
    graph, _, _ = engine.knn_graph(cd, k, device=dev)
    co = engine.spatial_order(cd, device=dev)
    gs = engine.relabel_graph(graph, co)
    std = engine.zscore_dense(X, rows=co.order)
    _, _, lag, _ = engine.lag_moran(gs, std.Z, g)
    sampler = ClockSampler(gpu_index)
    sampler.start()
    ms_tc = _timed_ms(torch, lambda: engine.lee_gemm(std.Z, lag, g, impl=2), reps=7)
    ms_f64 = _timed_ms(torch, lambda: engine.lee_gemm(std.Z, lag, g, impl=1), reps=3)
    ms_lag = _timed_ms(torch, lambda: engine.lag_moran(gs, std.Z, g), reps=5)
    clocks = sampler.stop()
    # accuracy of the tensor-core path against an FP64 evaluation of 4 096 sampled entries
    L = engine.lee_gemm(std.Z, lag, g, impl=2)
    rng = np.random.default_rng(0)
    ii = torch.from_numpy(rng.integers(0, g, 4096)).to(dev)
    jj = torch.from_numpy(rng.integers(0, g, 4096)).to(dev)
    ref = torch.empty(4096, dtype=torch.float64, device=dev)
    for c0 in range(0, 4096, 512):
        ref[c0:c0 + 512] = (std.Z[:, ii[c0:c0 + 512]].double() * lag[:, jj[c0:c0 + 512]].double()).sum(0)
    got = L[ii, jj].double()
    rel = ((got - ref).abs() / ref.abs().clamp_min(1e-300)).cpu().numpy()
    scale = float(ref.abs().max())
    # end to end through the public API on host arrays
    Xh = X.cpu().numpy()
    del X
    adata = AnnDataLite(Xh, obsm={"spatial": coords}, var_names=[f"g{i}" for i in range(g)])
    spatial.lees_l_matrix(adata, n_neighbors=k, impl=2, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    spatial.lees_l_matrix(adata, n_neighbors=k, impl=2, device=dev)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # CPU arm: the reference's per-pair loop [R autocorrelation.py:307-315, 1113-1155] as restated in the oracle,
    # on a sample of pairs without permutations, extrapolated to all G(G-1)/2 pairs
    cpu = None
    try:
        from oracle import restate

        Wc = graph.to_scipy("weights", np.float32)
        Zc = restate.zscore(Xh[:, :16])[0].astype(np.float32)
        pairs = [(a, b) for a in range(4) for b in range(4, 8)]
        rng_c = np.random.default_rng(0)
        restate.lees_l_pair(Zc[:, 0], Zc[:, 1], Wc, 0, rng_c)
        t0 = time.perf_counter()
        for a, b in pairs:
            restate.lees_l_pair(Zc[:, a], Zc[:, b], Wc, 0, rng_c)
        per_pair = (time.perf_counter() - t0) / len(pairs)
        cpu = {"seconds_per_pair": round(per_pair, 5), "pairs_sampled": len(pairs), "cores": 1, "kind": "port",
               "all_pairs_extrapolated_s": round(per_pair * g * (g - 1) / 2, 1),
               "sample": "oracle/restate.py lees_l_pair (scipy CSR x vector + dot, FP32, no permutations) on 16 of the 499 500 gene pairs"}
    except Exception as exc:
        cpu = {"error": f"{type(exc).__name__}: {exc}"}
    peaks = _peaks()
    tf32_peak = float(peaks.get("bf16_tflops", 1630.8)) / 2.0
    useful = 2.0 * n * g * g / (ms_tc / 1e3) / 1e12
    return {"workload": "C3: Lee's L over all gene pairs, 200k cells x 1000 genes, kNN k=6", "ms": round(ms_tc, 3),
            "useful_tflops": round(useful, 1), "issued_tf32_tflops": round(3.0 * useful, 1), "impl": "tcgen05 kind::tf32 cta_group::2 (CTA pair, 256 x 256 tile), 3xTF32 (impl=2)",
            "roofline": {"bound": "tensor", "achieved": round(3.0 * useful, 1), "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                         "frac": round(3.0 * useful / tf32_peak, 4), "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (dense TF32, burst)"},
            "fp64_exact_ms": round(ms_f64, 3), "lag_ms": round(ms_lag, 3),
            "rel_err_vs_fp64": {"p50": float(np.percentile(rel, 50)), "p99": float(np.percentile(rel, 99)), "max": float(rel.max()),
                                "max_abs_over_max_L": float((got - ref).abs().max()) / scale, "entries": 4096},
            "e2e": {"seconds": round(e2e_s, 3), "api": "spatialcore_b200.spatial.lees_l_matrix(adata[numpy host]) -> 1000 x 1000 DataFrame",
                    "h2d_bytes": int(Xh.nbytes + coords.nbytes), "d2h_bytes": int(4 * g * g)},
            "cpu_baseline": cpu, "clocks": clocks}


def nbhd_leg(spatial, AnnDataLite, engine, synthetic, dev, gpu_index):
    """BASELINE config 5: neighbourhood composition, kNN k=30 cell-type counts for 2M cells x 30 types."""
    import pandas as pd
    import torch

    n, k, T = 2_000_000, 30, 30
    coords = synthetic.coords_mixture(n, 2e4, 4)
    lab = synthetic.patchy_labels(coords, T, 5)
    cd = torch.from_numpy(coords).to(dev)
    ld = torch.from_numpy(lab).to(dev)

    def fused():
        _, _, prof = engine.knn_graph(cd, k, labels=ld, n_types=T, want_idx=False, device=dev)
        engine.profile_normalize(prof, True)

    sampler = ClockSampler(gpu_index)
    sampler.start()
    ms = _timed_ms(torch, fused, reps=5)
    ms_graph = _timed_ms(torch, lambda: engine.knn_graph(cd, k, device=dev), reps=5)
    clocks = sampler.stop()
    a = AnnDataLite(np.zeros((n, 1), np.float32), obsm={"spatial": coords})
    a.obs = pd.DataFrame({"ct": pd.Categorical.from_codes(lab, [f"t{i:02d}" for i in range(T)])})
    spatial.compute_neighborhood_profile(a, "ct", k=k, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    spatial.compute_neighborhood_profile(a, "ct", k=k, device=dev)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    cpu = None
    try:  # CPU arm: the reference's recipe (cKDTree.query(k + 1), drop self, count labels) restated in the oracle, on a sample
        from oracle import restate

        ns = 100_000
        t0 = time.perf_counter()
        restate.neighborhood_profile(coords[:ns], lab[:ns], T, k=k)
        dt = time.perf_counter() - t0
        cpu = {"seconds": round(dt, 2), "cells": ns, "cores": 1, "kind": "port", "extrapolated_2M_s": round(dt * n / ns, 1),
               "sample": "oracle/restate.py neighborhood_profile on the first 100 000 cells (vectorised numpy; the reference's Python "
                         "loops [R neighborhoods.py:223-233] measured 76 us per cell = ~150 s at 2 M in the survey)"}
    except Exception as exc:
        cpu = {"error": f"{type(exc).__name__}: {exc}"}
    nbytes = 17.0 * n + 4.0 * n * T
    peak = float(_peaks().get("hbm_gbs", 6650.0))
    return {"workload": "C5: neighbourhood composition, kNN k=30, 2M cells x 30 types", "ms": round(ms, 3), "cpu_baseline": cpu,
            "kernel": "kNN query with fused label-histogram epilogue (indices never materialised) + sc_profile_normalize",
            "roofline": {"bound": "hbm", "achieved": round(nbytes / (ms / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(nbytes / (ms / 1e3) / 1e9 / peak, 4), "bytes": nbytes,
                         "note": "algorithmic bytes 17N + 4NT (SURVEY.md §8d); the kernel is instruction-issue bound, not HBM bound"},
            "knn_graph_k30_2M_ms": round(ms_graph, 3),
            "e2e": {"seconds": round(e2e_s, 3), "api": "spatialcore_b200.spatial.compute_neighborhood_profile(adata[host], k=30)",
                    "h2d_bytes": int(coords.nbytes + n), "d2h_bytes": int(4 * n * T)},
            "clocks": clocks}


def local_moran_leg(spatial, AnnDataLite, synthetic, dev, gpu_index):
    """The only timing the reference publishes (docs/spatial/spatial_stats.md:202-215): batched
    ``local_morans_i`` on its 366 938-cell CosMx vignette, 5 / 10 / 20 genes, 10 permutations -- ~69 / 52 / 80 s on
    unstated hardware.  Same cell count, gene counts, k = 6 and permutation count on synthetic data, end to end through
    the public API on host arrays (six per-cell output matrices returned to the host)."""
    import torch

    n = 366_938
    coords = synthetic.coords_mixture(n, 8e3, 9)
    X = synthetic.expression_device(coords, 20, 9, device=dev).cpu().numpy()
    sampler = ClockSampler(gpu_index)
    sampler.start()
    out = {"workload": "local_morans_i, 366 938 cells, k=6, 10 permutations (the reference's published vignette timing)", "seconds": {},
           "reference_published_seconds": {"5": 69.0, "10": 52.0, "20": 80.0},
           "reference_hardware": "not stated (docs/spatial/spatial_stats.md:202-215; bar heights read off the figure)"}
    for g in (5, 10, 20):
        a = AnnDataLite(X[:, :g].copy(), obsm={"spatial": coords})
        spatial.local_morans_i(a, n_permutations=10, device=dev)  # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        spatial.local_morans_i(a, n_permutations=10, device=dev)
        torch.cuda.synchronize()
        out["seconds"][str(g)] = round(time.perf_counter() - t0, 4)
    out["vs_published"] = {k: round(out["reference_published_seconds"][k] / v, 1) for k, v in out["seconds"].items()}
    out["clocks"] = sampler.stop()
    return out


def knn_leg(engine, synthetic, dev, gpu_index):
    """kNN / radius graph build times (the 'kNN build ms' part of BASELINE.json's metric) on the configs' point sets."""
    import torch

    out = {}
    sampler = ClockSampler(gpu_index)
    sampler.start()
    peak = float(_peaks().get("hbm_gbs", 6650.0))
    for name, n, ext, k, gen, seed in (("C2_knn15_500k", 500_000, 1e4, 15, "mixture", 1), ("C5_knn30_2M", 2_000_000, 2e4, 30, "mixture", 4),
                                       ("knn15_5M", 5_000_000, 1.2e5, 15, "uniform", 3), ("knn6_5M", 5_000_000, 1.2e5, 6, "uniform", 3)):
        c = synthetic.coords_mixture(n, ext, seed) if gen == "mixture" else synthetic.coords_uniform(n, ext, seed)
        cd = torch.from_numpy(c).to(dev)
        ms = _timed_ms(torch, lambda: engine.knn_graph(cd, k, device=dev), reps=5)
        nbytes = 16.0 * n + 4.0 * n * k
        out[name] = {"ms": round(ms, 3), "gbs": round(nbytes / (ms / 1e3) / 1e9, 1), "frac_of_hbm_peak": round(nbytes / (ms / 1e3) / 1e9 / peak, 4)}
        del cd
    c = synthetic.coords_uniform(5_000_000, 1.2e5, 3)
    cd = torch.from_numpy(c).to(dev)
    r = synthetic.radius_for_mean_degree(5_000_000, 1.2e5, 20.0)
    ms = _timed_ms(torch, lambda: engine.radius_graph(cd, r, device=dev), reps=5)
    nnz = int(engine.radius_graph(cd, r, device=dev)[0].nnz)
    nbytes = 16.0 * 5_000_000 + 4.0 * nnz
    out["C4_radius_5M_deg20"] = {"ms": round(ms, 3), "nnz": nnz, "gbs": round(nbytes / (ms / 1e3) / 1e9, 1),
                                 "frac_of_hbm_peak": round(nbytes / (ms / 1e3) / 1e9 / peak, 4)}
    out["clocks"] = sampler.stop()
    try:  # CPU arm: the reference's own call [R autocorrelation.py:393-395] on the C2 point set
        from sklearn.neighbors import NearestNeighbors

        c2 = synthetic.coords_mixture(500_000, 1e4, 1)
        t0 = time.perf_counter()
        NearestNeighbors(n_neighbors=16, algorithm="ball_tree", n_jobs=-1).fit(c2).kneighbors(c2)
        out["cpu_baseline"] = {"C2_knn15_500k_s": round(time.perf_counter() - t0, 2), "cores": os.cpu_count(), "kind": "reference",
                               "sample": "sklearn NearestNeighbors(n_neighbors=k+1, algorithm='ball_tree').kneighbors, the reference's call, all cores"}
    except Exception as exc:
        out["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
    out["note"] = "exact kNN / radius graphs, FP64 distances, canonical CSR; algorithmic bytes 16N + 4*nnz; instruction-issue bound"
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from spatialcore_b200 import AnnDataLite, engine, spatial, synthetic
    from spatialcore_b200 import distributed as dist_util
    from spatialcore_b200.spatial import autocorrelation as ac

    import logging

    sc_log = logging.getLogger("spatialcore")  # the reference logger writes to stdout: keep stdout to the JSON line
    sc_log.setLevel(logging.ERROR)
    for h in sc_log.handlers:
        if isinstance(h, logging.StreamHandler):
            h.setStream(sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    n, g_total, P = w["n"], w["g"], w["perms"]
    # partition: permutations only.  Every rank holds all genes, so the gathered rows stay 4 KB wide and
    # the gather kernel keeps its single-GPU efficiency (gene blocks of 500 measured 75-86 % of peak
    # against 90 % at 1000); SC_BENCH_GENE_GROUPS=auto restores the hybrid gene-block x permutation grid.
    n_gene_groups = dist_util.gene_groups_for(g_total, world) if os.environ.get("SC_BENCH_GENE_GROUPS") == "auto" else 1
    n_perm_groups = world // n_gene_groups
    gi, pi = rank // n_perm_groups, rank % n_perm_groups
    g_lo, g_hi = dist_util.block_slice(g_total, gi, n_gene_groups)
    g = g_hi - g_lo
    perm_range = dist_util.block_slice(P, pi, n_perm_groups)
    group = None
    if world > 1:
        for j in range(n_gene_groups):  # every rank creates every sub-group, in the same order
            grp = dist.new_group(list(range(j * n_perm_groups, (j + 1) * n_perm_groups)))
            if j == gi:
                group = grp
    radius = workload_radius(w)

    coords = make_coords(w, args.seed)
    coords_dev = torch.from_numpy(coords).to(dev)
    X_dev = synthetic.expression_device(coords, g, seed=args.seed * 1000 + gi, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_results(I, p):
        """Per-gene results of every gene block -> full table on every rank (one all-gather)."""
        if world > 1:
            sizes = [dist_util.block_slice(g_total, j, n_gene_groups) for j in range(n_gene_groups)]
            pad = max(hi - lo for lo, hi in sizes)
            buf = torch.zeros(2, pad, dtype=torch.float64, device=dev)
            buf[0, : I.numel()] = I
            buf[1, : p.numel()] = p
            out = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(out, buf)
            firsts = [out[j * n_perm_groups] for j in range(n_gene_groups)]  # one rank per gene block
            I = torch.cat([o[0, : hi - lo] for o, (lo, hi) in zip(firsts, sizes)])
            p = torch.cat([o[1, : hi - lo] for o, (lo, hi) in zip(firsts, sizes)])
        return I.cpu().numpy(), p.cpu().numpy()

    # ---------------- device-resident leg -------------------------------------------------------
    for _ in range(args.warmup):
        I, p = device_step(engine, ac, coords_dev, X_dev, w, radius, args.seed, None, perm_range, group)
        gather_results(I, p)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = engine.launches()
    phase_ms = {}
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        events = []
        I, p = device_step(engine, ac, coords_dev, X_dev, w, radius, args.seed, events, perm_range, group)
        I_h, p_h = gather_results(I, p)
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(events[:-1], events[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + a.elapsed_time(b) / args.steps
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = engine.launches() - launches0
    ms = t_start.elapsed_time(t_end)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = g_total * P / (ms_per_step / 1e3)

    # roofline of the dominant kernel (perm_rows_kernel): algorithmic bytes per launch / launch time.
    # per gene-perm 4N(1+1/P) bytes; one launch covers PB permutations x g genes.
    PB = 16  # permutations per launch of the default kernel variant (bulk16)
    my_perms = perm_range[1] - perm_range[0]
    n_launch = (my_perms + PB - 1) // PB
    perm_ms = phase_ms.get("perms", float("nan"))
    bytes_per_launch = 4.0 * n * (1.0 + 1.0 / P) * g * (my_perms / n_launch)
    achieved = bytes_per_launch / (perm_ms / 1e3 / n_launch) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM bytes per launch of the same kernel at the same geometry, from the committed ncu --set full
    # capture (profiles/); None for geometries that were not captured
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r01b_perm_rows_bulk16_c4_ncu.json")))
        if args.workload == "C4" and g == 1000 and my_perms >= PB:
            traffic = cap["traffic_bytes_per_launch"] * (my_perms / n_launch) / PB
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "perm_rows_bulk_kernel<16> (cp.async.bulk gather pipeline, FP64 accumulate)", "achieved": round(achieved, 1), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of a full 16-permutation launch (profiles/r01b_perm_rows_bulk16_c4_ncu.json), scaled to the mean permutations per launch" if traffic else None,
                "launch_ms": round(perm_ms / n_launch, 4), "perms_per_launch": PB,
                "bytes_per_launch": bytes_per_launch}

    # the second kernel of the step: lag + Moran sums (shared-memory tile kernel; bound by the LDS crossbar, §8d)
    lag_ms = phase_ms.get("lag", float("nan")) + phase_ms.get("tiles", 0.0)
    nnz_edges = float(n) * (w["degree"] or w["k"])
    lag_bytes = 8.0 * n * g + 4.0 * nnz_edges + 4.0 * n
    lag_roofline = {"bound": "hbm", "kernel": "lag_tile_kernel (cp.async-staged neighbour unions in shared memory, FADD2) + tile build",
                    "achieved": round(lag_bytes / (lag_ms / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(lag_bytes / (lag_ms / 1e3) / 1e9 / peak, 4), "ms": round(lag_ms, 3),
                    "bytes": lag_bytes, "lds_gather_gbs": round(4.0 * nnz_edges * g / (lag_ms / 1e3) / 1e9, 1),
                    "note": "algorithmic bytes 8NG + 4*nnz + 4N (Z read, lag written, CSR); the kernel's own ceiling is the shared-memory "
                            "crossbar: 4*nnz*G gathered bytes at 128 B/clk/SM"}

    # ---------------- the other null, same workload (single-GPU runs) ---------------------------
    values_null = None
    if world == 1 and not args.no_values:
        try:
            values_null = values_null_leg(spatial, AnnDataLite, engine, coords, coords_dev, X_dev, w, radius, args.seed, local_rank)
        except Exception as exc:  # a secondary figure must never cost the headline line
            values_null = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    # ---------------- end-to-end legs through the public API, host buffers ------------------------
    # `e2e`: the drop-in defaults -- morans_i materialises obsp['spatial_connectivities'/'spatial_distances'] as the
    # reference always does [R autocorrelation.py:565-570]; `e2e_nograph`: the same call with write_graph=False.
    e2e = e2e_nograph = None
    if not args.no_e2e:
        pinned_slab = None
        if group is not None:
            # row-sharded ingest: a rank reads only its block of cells, so only that block is page-locked (below,
            # once the pageable buffer is filled: a copy into a partly registered range is rejected by the runtime)
            X_host = torch.empty((n, g), dtype=torch.float32)
        else:
            try:
                X_host = torch.empty((n, g), dtype=torch.float32, pin_memory=True)
            except RuntimeError:  # not enough lockable host memory for one pinned copy per rank
                X_host = torch.empty((n, g), dtype=torch.float32)
        X_host.copy_(X_dev)
        if group is not None:
            _, r_lo, r_hi = dist_util.row_block(n, dist.get_rank(group), dist.get_world_size(group))
            torch.cuda.synchronize()
            rc = torch.cuda.cudart().cudaHostRegister(X_host[r_lo:r_hi].data_ptr(), (r_hi - r_lo) * g * 4, 0)
            pinned_slab = (X_host[r_lo:r_hi].data_ptr(), int(rc))
        del X_dev
        torch.cuda.empty_cache()
        coords_host = torch.empty((n, 2), dtype=torch.float64, pin_memory=True)
        coords_host.copy_(torch.from_numpy(coords))
        Xn, cn = X_host.numpy(), coords_host.numpy()
        names = [f"g{i}" for i in range(g_lo, g_hi)]
        graph_bytes = [0]

        def e2e_step(write_graph):
            adata = AnnDataLite(Xn, obsm={"spatial": cn}, var_names=names)
            spatial.morans_i(adata, n_neighbors=w["k"] or 6, n_permutations=P, seed=args.seed, radius=radius,
                             perm_source="philox", write_graph=write_graph, shard="perms" if group is not None else "none",
                             ingest="sharded" if group is not None else "replicated", group=group, device=dev)
            df = adata.uns["morans_i"]
            if write_graph and "spatial_connectivities" in adata.obsp:
                c_, d_ = adata.obsp["spatial_connectivities"], adata.obsp["spatial_distances"]
                graph_bytes[0] = int(c_.indices.nbytes + c_.indptr.nbytes + d_.data.nbytes)
            return (torch.from_numpy(df["I"].to_numpy(copy=True)).to(dev),
                    torch.from_numpy(df["p_value"].to_numpy(copy=True)).to(dev))

        def e2e_run(write_graph, warm, steps):
            for _ in range(warm):
                gather_results(*e2e_step(write_graph))
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                gather_results(*e2e_step(write_graph))
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()) / steps

        h2d = int(Xn.nbytes * (n_gene_groups if group is not None else world) + cn.nbytes * world)
        note = "bytes per step over all ranks" + ("; row-sharded ingest: each rank uploads N/W cells; fused standardise + all-gather + re-order kernel over NVLink peer memory" if group is not None else "")
        api = "spatialcore_b200.spatial.morans_i(adata[numpy, pinned host], shard='perms', ingest='sharded') per rank"
        if group is not None:
            os.environ["SC_INGEST_PROFILE"] = "1"  # CUDA-event breakdown of the row-sharded ingest (one sync per call)
        s_graph = e2e_run(True, max(1, min(args.warmup, 2)), args.steps)
        e2e = {"value": round(g_total * P / s_graph, 1), "unit": UNIT, "ms_per_step": round(s_graph * 1e3, 2),
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(2 * 8 * g + 3 * 8 + graph_bytes[0]),
               "h2d_note": note + ("" if pinned_slab is None else f"; host block page-locked: {pinned_slab[1] == 0}"), "api": api + ", drop-in defaults (write_graph=True: obsp graph slots materialised "
               + ("on rank 0 of the group" if group is not None else "as the reference does") + ", host assembly overlapped with the permutation kernels)"}
        if group is not None:
            e2e["ingest_ms"] = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in ac.last_ingest_ms.items()}
        s_lean = e2e_run(False, 1, max(1, min(args.steps, 2)))
        e2e_nograph = {"value": round(g_total * P / s_lean, 1), "unit": UNIT, "ms_per_step": round(s_lean * 1e3, 2),
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(2 * 8 * g + 3 * 8), "api": api + ", write_graph=False"}
        if pinned_slab is not None and pinned_slab[1] == 0:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(pinned_slab[0])
        del X_host, Xn
        X_dev = None

    # ---------------- the other BASELINE configs (single-GPU runs): Lee's L, neighbourhoods, kNN ------
    legs = {}
    if world == 1 and not args.no_legs:
        torch.cuda.empty_cache()
        for name, fn in (("lee", lambda: lee_leg(spatial, AnnDataLite, engine, synthetic, dev, local_rank)),
                         ("nbhd", lambda: nbhd_leg(spatial, AnnDataLite, engine, synthetic, dev, local_rank)),
                         ("knn", lambda: knn_leg(engine, synthetic, dev, local_rank)),
                         ("local_moran", lambda: local_moran_leg(spatial, AnnDataLite, synthetic, dev, local_rank))):
            try:
                legs[name] = fn()
            except Exception as exc:  # failure-isolated: a leg can never cost the headline line
                legs[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    # ---------------- CPU baseline on a bounded sample (rank 0, single-GPU runs only) --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_sample(w, coords, radius, args.seed, values=isinstance(values_null, dict) and "error" not in values_null)
        if "values_null" in cpu:  # the CPU arm of the second null sits with its GPU figure
            values_null["cpu_baseline"] = cpu.pop("values_null")

    if rank == 0:
        config = workload_config(args.workload, w, radius)
        line = {
            "metric": metric_name(args.workload), "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 storage, f64 accumulation", "data": "synthetic",
            "config": config,
            "arm": {"perm_source": "philox (on-device bijection)",
                    "sharding": f"{n_gene_groups} gene block(s) x {n_perm_groups} permutation group(s) over {world} rank(s)"},
            "phases_ms": {k: round(v, 3) for k, v in phase_ms.items()},
            "knn_build_ms": round(phase_ms.get("graph", float("nan")), 3),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "lag_roofline": lag_roofline,
            "e2e": e2e, "e2e_nograph": e2e_nograph, "cpu_baseline": cpu,
            "values_null": values_null, "lee": legs.get("lee"), "nbhd": legs.get("nbhd"), "knn": legs.get("knn"),
            "local_moran": legs.get("local_moran"),
            "check": {"I_mean": float(np.nanmean(I_h)), "p_min": float(np.nanmin(p_h)), "n_sig_0.01": int((p_h <= 0.01).sum())},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path (sklearn graph + the squidpy/scanpy Moran loop, ported in
# oracle/moran_port.c because squidpy is not installable here) on a bounded sample
# ------------------------------------------------------------------------------------------------


def cpu_graph(w, coords, radius):
    """Neighbour graph the way the reference path builds it (squidpy -> sklearn NearestNeighbors)."""
    from scipy import sparse
    from sklearn.neighbors import NearestNeighbors

    t0 = time.perf_counter()
    if radius is not None:
        nn = NearestNeighbors(radius=radius, n_jobs=-1).fit(coords)
        adj = nn.radius_neighbors_graph(mode="connectivity").tocsr()
    else:
        nn = NearestNeighbors(n_neighbors=w["k"], n_jobs=-1).fit(coords)
        adj = nn.kneighbors_graph(mode="connectivity").tocsr()
    dt = time.perf_counter() - t0
    adj = adj.astype(np.float64)
    rs = np.asarray(adj.sum(axis=1)).ravel()
    rs[rs == 0] = 1.0
    adj.data /= np.repeat(rs, np.diff(adj.indptr))
    return sparse.csr_matrix(adj), dt


def cpu_expression(coords, g, seed):
    rng = np.random.default_rng(seed)
    return np.log1p(rng.poisson(0.4, (g, coords.shape[0]))).astype(np.float64)  # (G, N) like squidpy's vals


def cpu_sample_plan(w):
    """Genes x permutations of the bounded CPU sample: ~1.5e10 gather-FMAs (10-30 s on 8-64 cores)."""
    nnz = w["n"] * (w["degree"] or w["k"])
    budget = 1.5e10
    gene_perms = max(4, int(budget / nnz))
    g = int(min(w["g"], max(4, min(64, gene_perms // 2))))
    p = int(max(1, min(w["perms"], gene_perms // g)))
    return g, p


def cpu_values_null_sample(w, graph, seed):
    """CPU arm of the ``values_null`` leg: the permutation loop of the reference's own ``local_morans_i``
    [R spatial/autocorrelation.py:877-884] -- ``Zs = Z[perm]; lag = W @ Zs; I = Zs * lag`` with scipy's
    CSR x dense product in FP32 -- as restated in oracle/restate.py, on a bounded sample (full N, a few
    genes, one or two permutations).  scipy's product is single-threaded, as in the reference."""
    from oracle import restate

    n = w["n"]
    nnz = graph.nnz
    g_cpu = int(max(2, min(w["g"], 8, 4e9 / max(nnz, 1))))
    p_cpu = 2 if nnz * g_cpu <= 2e9 else 1
    rng = np.random.default_rng(seed + 17)
    Z = rng.standard_normal((n, g_cpu), dtype=np.float32)
    W = graph.astype(np.float32)
    perms = restate.squidpy_perm_indices(n, p_cpu, seed + 17)  # default_rng(seed).permutation(n) per permutation
    restate.morans_values_null(Z[:, :1], W, perms[:1])  # touch pages
    t0 = time.perf_counter()
    restate.morans_values_null(Z, W, perms)
    dt = time.perf_counter() - t0
    return {"value": round(g_cpu * p_cpu / dt, 2), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"full N={n} cells, {g_cpu} genes x {p_cpu} permutation(s), oracle/restate.py morans_values_null "
                      f"(the reference's own scipy CSR x dense loop, FP32)",
            "seconds": round(dt, 2)}


def cpu_baseline_sample(w, coords, radius, seed, graph=None, values=False, graph_s=None):
    """The reference's CPU path on a bounded sample: full N, ``g_cpu`` genes, ``p_cpu`` permutations.  One call
    of the port executes ``p_cpu + 1`` passes over the graph per gene (the observed statistic plus every
    permutation, as squidpy does), so the rate is ``g_cpu * (p_cpu + 1) / seconds`` gene-passes per second --
    the unit the GPU arm counts (one gene-permutation = one pass).  ``value_incl_graph`` extrapolates the
    whole job (graph build once + G * (P + 1) passes) and expresses it in the GPU arm's terms, G * P / seconds."""
    from oracle import port, restate

    port.use_all_cores()
    g_cpu, p_cpu = cpu_sample_plan(w)
    n = w["n"]
    if graph is None:
        graph, graph_s = cpu_graph(w, coords, radius)
    vals = cpu_expression(coords, g_cpu, seed)
    perms = restate.squidpy_perm_indices(n, p_cpu, seed).astype(np.int32)
    port.morans_i(graph, vals[:1, : n], None)  # touch pages / thread pool
    t0 = time.perf_counter()
    port.morans_i(graph, vals, perms)
    dt = time.perf_counter() - t0
    rate = g_cpu * (p_cpu + 1) / dt
    out = {"value": round(rate, 2), "unit": UNIT, "cores": port.threads(), "kind": "port",
           "sample": f"full N={n} cells, {g_cpu} genes x ({p_cpu} permutations + the observed pass) = {g_cpu * (p_cpu + 1)} gene-passes, "
                     f"oracle/moran_port.c (OpenMP over genes, CSR row gather per permutation)",
           "seconds": round(dt, 2), "graph_build_s": None if graph_s is None else round(graph_s, 2),
           "graph_build": "sklearn NearestNeighbors (squidpy's call), n_jobs=-1"}
    if graph_s is not None:
        G, P = w["g"], w["perms"]
        full_job_s = graph_s + G * (P + 1) / rate
        out["value_incl_graph"] = round(G * P / full_job_s, 2)
        out["value_incl_graph_note"] = (f"whole job extrapolated: graph build {graph_s:.1f} s + {G} x ({P} + 1) gene-passes at the sampled rate "
                                        f"= {full_job_s:.0f} s; expressed as G*P/seconds like the GPU arm's value")
    if values:
        try:
            out["values_null"] = cpu_values_null_sample(w, graph, seed)
        except Exception as exc:  # secondary figure
            out["values_null"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import port

    w = WORKLOADS[args.workload]
    coords = make_coords(w, args.seed)
    radius = workload_radius(w)
    graph, graph_s = cpu_graph(w, coords, radius)
    samples = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline_sample(w, coords, radius, args.seed + i, graph=graph, graph_s=graph_s)
        if i >= args.warmup:
            samples.append(r)
    value = float(np.mean([s["value"] for s in samples]))
    secs = float(np.mean([s["seconds"] for s in samples]))
    incl = float(np.mean([s["value_incl_graph"] for s in samples]))
    cpu = dict(samples[-1], value=round(value, 2), value_incl_graph=round(incl, 2), graph_build_s=round(graph_s, 2), kind="port")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": round(incl, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, w, radius),
        "arm": {"note": "CPU reference path: sklearn graph build (timed once, graph_build_s) + the squidpy/scanpy "
                        "Moran permutation loop as ported in oracle/moran_port.c (squidpy itself is not installable "
                        "offline); each step is a bounded sample of the workload (full N, a gene subset, a few permutations); "
                        "`value` is the whole-job figure including the graph build (cpu_baseline.value_incl_graph), like "
                        "the b200 arm's; cpu_baseline.value is the permutation loop alone"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(incl, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cores": port.threads(),
    }
    print(json.dumps(line))


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the single JSON line
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
