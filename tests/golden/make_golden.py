"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference (via oracle/ref_shim.py).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/restate.py) and the CUDA path to outputs of the
reference's own code on seeded inputs.  Input recipes are regenerated in the tests from the
seeds recorded here (``inputs.py``), so only outputs are stored.
"""

import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from tests.golden import inputs  # noqa: E402


def main() -> None:
    ac, nb, AD = ref_shim.load()

    # ---------------- G0: SURVEY Appendix B recipe --------------------------------------
    coords, X32 = inputs.g0()
    X64 = X32.astype(np.float64)
    names = [f"g{i}" for i in range(X32.shape[1])]

    out = {}
    for k in (6, 15):
        W = ac.build_spatial_weights(AD(X32, obsm={"spatial": coords}), n_neighbors=k)
        assert W.has_sorted_indices
        out[f"W_k{k}_indices"] = W.indices.astype(np.int32)
        out[f"W_k{k}_indptr"] = W.indptr.astype(np.int32)
        out[f"W_k{k}_data0"] = W.data[:8].copy()
    Ws = ac.build_spatial_weights(AD(X32, obsm={"spatial": coords}), n_neighbors=6, include_self=True)
    out["W_k6_self_indices"] = Ws.indices.astype(np.int32)
    out["W_k6_self_data0"] = Ws.data[:8].copy()

    pairs = [("g0", "g1"), ("g1", "g0"), ("g0", "g0"), ("g3", "g7")]
    for tag, X in (("f32", X32), ("f64", X64)):
        r = ac.lees_l(AD(X, obsm={"spatial": coords}, var_names=names), pairs, n_neighbors=6, n_permutations=99, seed=0)
        out[f"lee_{tag}_L"] = np.array([d["L"] for d in r], dtype=np.float64)
        out[f"lee_{tag}_p"] = np.array([d["p_value"] for d in r], dtype=np.float64)

    a = AD(X64, obsm={"spatial": coords}, var_names=names)
    ac.local_morans_i(a, genes=["g0", "g1", "g2"], n_neighbors=6, n_permutations=99, seed=0)
    for s in ("I", "z", "lag", "p", "p_adj", "quadrant"):
        out[f"lm_{s}"] = np.asarray(a.obsm[f"local_morans_{s}"])
    a = AD(X64, obsm={"spatial": coords}, var_names=names)
    ac.local_morans_i(a, genes=["g4", "g5"], n_neighbors=6, n_permutations=19, seed=3, fdr_correction="bonferroni", alpha=0.5)
    for s in ("p", "p_adj", "quadrant"):
        out[f"lm2_{s}"] = np.asarray(a.obsm[f"local_morans_{s}"])

    a = AD(X64, obsm={"spatial": coords}, var_names=names)
    ac.lees_l_local(a, gene_pairs=[("g0", "g1"), ("g2", "g3")], n_neighbors=6, n_permutations=19,
                    compute_cell_pvalues=True, significance_filter=True, alpha=0.2, seed=0)
    for key in ("g0_g1", "g2_g3"):
        out[f"ll_{key}_L"] = a.obs[f"{key}_lees_l"].to_numpy()
        out[f"ll_{key}_p"] = a.obs[f"{key}_pvalue"].to_numpy()
        out[f"ll_{key}_q"] = a.obs[f"{key}_quadrant"].cat.codes.to_numpy().astype(np.int8)
        prm = a.uns[f"{key}_lees_l_params"]
        out[f"ll_{key}_global"] = np.array([prm["global_L"], prm["global_pvalue"]])
    # continuous expression: tie-free per-cell p-values (see inputs.g0_continuous)
    coords_c, Xc = inputs.g0_continuous()
    namesc = [f"g{i}" for i in range(Xc.shape[1])]
    a = AD(Xc, obsm={"spatial": coords_c}, var_names=namesc)
    ac.local_morans_i(a, genes=["g0", "g1", "g2"], n_neighbors=6, n_permutations=99, seed=0)
    for s in ("I", "z", "lag", "p", "p_adj", "quadrant"):
        out[f"lmc_{s}"] = np.asarray(a.obsm[f"local_morans_{s}"])
    a = AD(Xc, obsm={"spatial": coords_c}, var_names=namesc)
    ac.lees_l_local(a, gene_pairs=[("g0", "g1"), ("g2", "g3")], n_neighbors=6, n_permutations=19,
                    compute_cell_pvalues=True, significance_filter=True, alpha=0.2, seed=0)
    for key in ("g0_g1", "g2_g3"):
        out[f"llc_{key}_L"] = a.obs[f"{key}_lees_l"].to_numpy()
        out[f"llc_{key}_p"] = a.obs[f"{key}_pvalue"].to_numpy()
        out[f"llc_{key}_q"] = a.obs[f"{key}_quadrant"].cat.codes.to_numpy().astype(np.int8)
        prm = a.uns[f"{key}_lees_l_params"]
        out[f"llc_{key}_global"] = np.array([prm["global_L"], prm["global_pvalue"]])
    r = ac.lees_l(AD(Xc, obsm={"spatial": coords_c}, var_names=namesc),
                  [("g0", "g1"), ("g1", "g0"), ("g5", "g5")], n_neighbors=6, n_permutations=99, seed=0)
    out["leec_L"] = np.array([d["L"] for d in r])
    out["leec_p"] = np.array([d["p_value"] for d in r])
    np.savez_compressed(os.path.join(HERE, "ref_g0.npz"), **out)
    print("ref_g0.npz", {k: v.shape for k, v in out.items()})

    # ---------------- neighbourhood profiles --------------------------------------------
    out = {}
    coords, labels = inputs.nbhd()
    obs = pd.DataFrame({"ct": pd.Categorical([f"t{c:02d}" for c in labels])})
    for k in (5, 30):
        a = AD(np.zeros((coords.shape[0], 1), np.float32), obs=obs.copy(), obsm={"spatial": coords})
        nb.compute_neighborhood_profile(a, "ct", method="knn", k=k)
        out[f"knn{k}_norm"] = a.obsm["neighborhood_profile"]
        out["celltypes"] = np.array(a.uns["neighborhood_profile_celltypes"])
    a = AD(np.zeros((coords.shape[0], 1), np.float32), obs=obs.copy(), obsm={"spatial": coords})
    nb.compute_neighborhood_profile(a, "ct", method="knn", k=30, normalize=False)
    out["knn30_raw"] = a.obsm["neighborhood_profile"]
    for r in (inputs.NBHD_RADIUS,):
        a = AD(np.zeros((coords.shape[0], 1), np.float32), obs=obs.copy(), obsm={"spatial": coords})
        nb.compute_neighborhood_profile(a, "ct", method="radius", radius=r, normalize=False)
        out["radius_raw"] = a.obsm["neighborhood_profile"]
        a = AD(np.zeros((coords.shape[0], 1), np.float32), obs=obs.copy(), obsm={"spatial": coords})
        nb.compute_neighborhood_profile(a, "ct", method="radius", radius=r, normalize=True)
        out["radius_norm"] = a.obsm["neighborhood_profile"]
    np.savez_compressed(os.path.join(HERE, "ref_nbhd.npz"), **out)
    print("ref_nbhd.npz", {k: v.shape for k, v in out.items()})

    # ---------------- kNN on clustered / larger-k inputs --------------------------------
    out = {}
    coords = inputs.clustered(4000, seed=11)
    for k in (6, 15, 30, 50):
        W = ac.build_spatial_weights(AD(np.zeros((4000, 1), np.float32), obsm={"spatial": coords}), n_neighbors=k)
        out[f"W_k{k}_indices"] = W.indices.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_knn_clustered.npz"), **out)
    print("ref_knn_clustered.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
