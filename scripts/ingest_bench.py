"""Row-sharded ingest at C4 size under torchrun: fused peer-memory scatter vs NCCL all-gather + re-order.
Prints, on rank 0, the max-over-ranks wall time of host matrix -> standardised Z in spatial order on every GPU."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import AnnDataLite, engine, synthetic
from spatialcore_b200.spatial import autocorrelation as ac
import logging; logging.getLogger("spatialcore").setLevel(logging.ERROR)

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
from spatialcore_b200 import distributed as du
bound = du.bind_host_to_gpu(local) if os.environ.get("SC_BIND_NUMA", "1") == "1" else False
dist.init_process_group("nccl", device_id=dev)
n, g = 5_000_000, 1000
coords = synthetic.coords_uniform(n, 1.2e5, 3)
Xd = synthetic.expression_device(coords, g, seed=3000, device=dev)
X = torch.empty((n, g), dtype=torch.float32, pin_memory=True); X.copy_(Xd); del Xd
a = AnnDataLite(X.numpy(), obsm={"spatial": coords}, var_names=[f"g{i}" for i in range(g)])
co = engine.spatial_order(torch.from_numpy(coords).to(dev))
res = {}
for name, env in (("fused_peer_memory", None), ("nccl_allgather_reorder", "1")):
    if env: os.environ["SC_INGEST_NCCL"] = env
    else: os.environ.pop("SC_INGEST_NCCL", None)
    ts = []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        std = ac._standardize_row_sharded(a, None, list(a.var_names), dev, co, None)
        torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it > 0: ts.append(float(dt.item()))
        chk = float(std.Z[:, :8].double().abs().sum().item())
        del std
    res[name] = {"ms": round(1e3 * float(np.median(ts)), 2), "checksum": chk}
if rank == 0:
    print({"world": world, "n": n, "g": g, "numa_bound": bound, "cpus": len(os.sched_getaffinity(0)), **res}, flush=True)
dist.destroy_process_group()
