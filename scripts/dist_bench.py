"""torchrun timing (NCCL) of the two other sharded entry points: local_morans_i (gene batches over ranks, six
per-cell matrices all-gathered) at C2's cell count and lees_l_matrix (permutations over ranks, one all-reduce of
the G x G exceedance counts) at C3's cell count.  Prints one JSON object on rank 0; a single-rank run of the same
call (shard="none") on every rank gives the baseline."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import AnnDataLite, spatial, synthetic
import logging; logging.getLogger("spatialcore").setLevel(logging.ERROR)

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

def timed(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

out = {"world": world}
# local Moran: 500k cells x 48 genes, batch_size 6 -> 8 batches, P = 19 (host-bound: the six per-cell matrices) and 199
n, g = 500_000, 48
c = synthetic.coords_mixture(n, 1e4, 1)
X = synthetic.expression_device(c, g, 1, device=dev).cpu().numpy()
for P in (19, 199):
    for mode in ("none", "genes"):
        a = AnnDataLite(X, obsm={"spatial": c})
        out[f"local_morans_500k_x48_P{P}_shard_{mode}_s"] = round(timed(lambda: spatial.local_morans_i(a, n_neighbors=15, n_permutations=P, batch_size=6, perm_source="philox", shard=mode, device=dev), reps=1 if P > 100 else 2), 3)
out["local_morans_allgather_bytes"] = int(21 * n * g)  # five float32 matrices + one int8 matrix, every rank receives all of them
# Lee all-pairs permutation p-values: 200k cells x 1000 genes, 64 permutations
n, g, P = 200_000, 1000, 64
c = synthetic.coords_mixture(n, 6e3, 2)
X = synthetic.expression_device(c, g, 2, device=dev).cpu().numpy()
for mode in ("none", "perms"):
    a = AnnDataLite(X, obsm={"spatial": c})
    out[f"lees_l_matrix_200k_x1000_P64_shard_{mode}_s"] = round(timed(lambda: spatial.lees_l_matrix(a, n_permutations=P, seed=1, perm_source="philox", shard=mode, device=dev)), 3)
if rank == 0:
    for k in ("local_morans_500k_x48_P19", "local_morans_500k_x48_P199", "lees_l_matrix_200k_x1000_P64"):
        base = out[[x for x in out if x.startswith(k + "_shard") and x.endswith("none_s")][0]]
        sh = out[[x for x in out if x.startswith(k + "_shard") and not x.endswith("none_s")][0]]
        out[k + "_speedup"] = round(base / sh, 2)
        out[k + "_efficiency"] = round(base / sh / world, 3)
    print(json.dumps(out))
dist.destroy_process_group()
