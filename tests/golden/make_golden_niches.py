"""Freeze the UNMODIFIED reference's ``identify_niches`` output (this container only):

    python tests/golden/make_golden_niches.py

Input: the frozen reference profile ``ref_nbhd.npz['knn30_norm']`` (6000 cells x 8 types) and a
harder mixture-of-Dirichlet profile generated here.  Output: ``ref_niches.npz``.
"""
import os
import sys

import numpy as np


HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from tests.golden import inputs  # noqa: E402


def main() -> None:
    _, nb, AD = ref_shim.load()
    out = {}
    prof = np.load(os.path.join(HERE, "ref_nbhd.npz"))["knn30_norm"]
    for tag, P, k in (("nbhd", prof, 5), ("dirichlet", inputs.niche_profiles(), 6)):
        a = AD(np.zeros((P.shape[0], 1), np.float32), obsm={})
        a.obsm["neighborhood_profile"] = P
        nb.identify_niches(a, n_niches=k, random_state=0)
        out[f"{tag}_labels"] = a.obs["niche"].cat.codes.to_numpy().astype(np.int32)
        out[f"{tag}_centroids"] = np.asarray(a.uns["niche_centroids"])
        out[f"{tag}_inertia"] = np.float64(a.uns["niche_params"]["inertia"])
        out[f"{tag}_k"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "ref_niches.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
