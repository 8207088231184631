"""torchrun check (NCCL): morans_i with shard='genes' and shard='perms' on W ranks returns, on every
rank, the same table as a single-rank run (Philox permutations are addressed by global index); the same
for the row-sharded ingest, local_morans_i (gene batches) and lees_l_matrix (permutations)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import AnnDataLite, spatial

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import logging; logging.getLogger("spatialcore").setLevel(logging.ERROR)
rng = np.random.default_rng(0)
n, g = 20000, 37
coords = rng.uniform(0, 1000, (n, 2))
X = np.log1p(rng.poisson(0.7, (n, g))).astype(np.float32) + 0.05 * rng.normal(size=(n, g)).astype(np.float32)
dev = torch.device("cuda", local)
ref = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=101, seed=4, perm_source="philox",
                       shard="none", device=dev).uns["morans_i"]
for mode in ("genes", "perms"):
    got = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=101, seed=4, perm_source="philox",
                           shard=mode, device=dev).uns["morans_i"]
    assert got["gene"].tolist() == ref["gene"].tolist()
    assert np.array_equal(got["p_value"].to_numpy(), ref["p_value"].to_numpy()), mode
    np.testing.assert_allclose(got["I"].to_numpy(), ref["I"].to_numpy(), rtol=1e-12)
    np.testing.assert_allclose(got["z_score"].to_numpy(), ref["z_score"].to_numpy(), rtol=1e-12)
    if rank == 0: print("shard=%s ok on %d ranks" % (mode, dist.get_world_size()), flush=True)
# row-sharded ingest: pooled moments differ in the last FP64 bit -> Z agrees to FP32 rounding
sh = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=101, seed=4, perm_source="philox",
                      shard="perms", ingest="sharded", device=dev).uns["morans_i"]
np.testing.assert_allclose(sh["I"].to_numpy(), ref["I"].to_numpy(), rtol=1e-6, atol=1e-9)
assert (sh["p_value"].to_numpy() != ref["p_value"].to_numpy()).mean() <= 0.03
Xc = X.copy(); Xc[:, 5] = 0.75  # zero-variance gene must pool to exactly zero variance
shc = spatial.morans_i(AnnDataLite(Xc, obsm={"spatial": coords}), n_permutations=9, seed=4, perm_source="philox",
                       shard="perms", ingest="sharded", device=dev).uns["morans_i"]
assert np.isnan(shc["I"].to_numpy()[5]) and np.isfinite(np.delete(shc["I"].to_numpy(), 5)).all()
os.environ["SC_INGEST_NCCL"] = "1"  # the NCCL all-gather variant must give the same Z bit for bit
sh2 = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=101, seed=4, perm_source="philox",
                       shard="perms", ingest="sharded", device=dev).uns["morans_i"]
del os.environ["SC_INGEST_NCCL"]
assert np.array_equal(sh2["I"].to_numpy(), sh["I"].to_numpy()) and np.array_equal(sh2["p_value"].to_numpy(), sh["p_value"].to_numpy())
sh3 = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=101, seed=4, perm_source="philox",
                       shard="perms", ingest="sharded", device=dev).uns["morans_i"]  # cached symmetric buffer re-used
assert np.array_equal(sh3["I"].to_numpy(), sh["I"].to_numpy())
if rank == 0: print("ingest=sharded ok (fused peer-memory scatter == NCCL all-gather)", flush=True)
rep = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=33, seed=4, perm_source="replay",
                       shard="perms", device=dev).uns["morans_i"]
one = spatial.morans_i(AnnDataLite(X, obsm={"spatial": coords}), n_permutations=33, seed=4, perm_source="replay",
                       shard="none", device=dev).uns["morans_i"]
assert np.array_equal(rep["p_value"].to_numpy(), one["p_value"].to_numpy())
if rank == 0: print("replay shard=perms ok", flush=True)
# local_morans_i (gene batches over ranks, per-cell matrices all-gathered) and lees_l_matrix (permutations
# over ranks): every rank returns the single-process result, replay and Philox mode
Xs, cs = X[:6000, :9], coords[:6000]
for src in ("philox", "replay"):
    one = AnnDataLite(Xs, obsm={"spatial": cs}); many = AnnDataLite(Xs, obsm={"spatial": cs})
    spatial.local_morans_i(one, n_permutations=19, seed=4, batch_size=2, perm_source=src, shard="none", device=dev)
    spatial.local_morans_i(many, n_permutations=19, seed=4, batch_size=2, perm_source=src, shard="genes", device=dev)
    for key in ("I", "z", "lag", "p", "p_adj", "quadrant"):
        assert np.array_equal(one.obsm[f"local_morans_{key}"], many.obsm[f"local_morans_{key}"]), (src, key)
    L1, p1 = spatial.lees_l_matrix(AnnDataLite(Xs, obsm={"spatial": cs}), n_permutations=13, seed=6, perm_source=src,
                                   shard="none", impl=1, device=dev)
    L2, p2 = spatial.lees_l_matrix(AnnDataLite(Xs, obsm={"spatial": cs}), n_permutations=13, seed=6, perm_source=src,
                                   shard="perms", impl=1, device=dev)
    assert np.array_equal(L1.to_numpy(), L2.to_numpy()) and np.array_equal(p1.to_numpy(), p2.to_numpy()), src
if rank == 0: print("local_morans_i shard=genes and lees_l_matrix shard=perms ok", flush=True)
dist.destroy_process_group()
