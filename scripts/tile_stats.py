"""Union / word-count statistics of the tile form at C4 (or C2): how many 256-row chunks exceed the shared-memory
budget and run through the direct-gather fallback.  Reads the first two arrays of the tile blob (ucount, wtotal)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialcore_b200 import engine as eng, synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "C4"
if which == "C4":
    n = 5_000_000; c = synthetic.coords_uniform(n, 1.2e5, 3); cd = torch.from_numpy(c).cuda()
    graph, _ = eng.radius_graph(cd, synthetic.radius_for_mean_degree(n, 1.2e5, 20.0))
else:
    n = 500_000; c = synthetic.coords_mixture(n, 1e4, 1); cd = torch.from_numpy(c).cuda()
    graph, _, _ = eng.knn_graph(cd, 15)
co = eng.spatial_order(cd)
gs = eng.relabel_graph(graph, co)
n_chunks = (n + 255) // 256
pad = (4 * n_chunks + 255) // 256 * 256
t = gs.tiles.cpu().numpy()
ucount = t[:4 * n_chunks].view(np.int32); wtotal = t[pad:pad + 4 * n_chunks].view(np.int32)
deg = (gs.indptr[1:] - gs.indptr[:-1]).cpu().numpy() if gs.indptr is not None else np.full(n, gs.k_fixed)
print(which, "chunks", n_chunks, "overflow", int((ucount < 0).sum()), f"({100 * (ucount < 0).mean():.2f} %)", "max degree", int(deg.max()))
ok = ucount[ucount >= 0]
print("union rows of the chunks that fit: mean %.1f p50 %d p99 %d max %d" % (ok.mean(), np.percentile(ok, 50), np.percentile(ok, 99), ok.max()))
print("words of the chunks that fit: mean %.1f p99 %d max %d" % (wtotal[ucount >= 0].mean(), np.percentile(wtotal[ucount >= 0], 99), wtotal[ucount >= 0].max()))
# exact unions of the overflow chunks (host): how large would the tile have to be?
ip = gs.indptr.cpu().numpy() if gs.indptr is not None else np.arange(0, (n + 1) * gs.k_fixed, gs.k_fixed)
ix = gs.indices.cpu().numpy().reshape(-1)
ov = np.nonzero(ucount < 0)[0]
sizes, words = [], []
for ch in ov[:4000]:
    r0, r1 = ch * 256, min(n, ch * 256 + 256)
    cols = ix[ip[r0]:ip[r1]]
    sizes.append(len(np.union1d(cols, np.arange(r0, r1))))
    d = np.diff(ip[r0:r1 + 1]); words.append(int(((d + 3) // 4 * 4).sum()))
if sizes:
    sizes = np.array(sizes); words = np.array(words)
    print("overflow chunks: union p50 %d p90 %d p99 %d max %d; words p50 %d max %d" % (np.percentile(sizes, 50), np.percentile(sizes, 90), np.percentile(sizes, 99), sizes.max(), np.percentile(words, 50), words.max()))
    for cap in (608, 640, 656, 704, 768):
        print("  a cap of", cap, "rows would leave", int((sizes > cap).sum()), "of", len(sizes))
