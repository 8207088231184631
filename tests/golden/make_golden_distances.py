"""Freeze the UNMODIFIED reference's ``calculate_domain_distances`` outputs (this container only):

    python tests/golden/make_golden_distances.py   ->  ref_distances.npz
"""
import importlib
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from tests.golden import inputs  # noqa: E402

CASES = [  # tag, source column, target column, metric, mode
    ("min_both", "src", "tgt", "minimum", "both"),
    ("min_matrix", "src", "tgt", "minimum", "matrix"),
    ("centroid_both", "src", "tgt", "centroid", "both"),
    ("mean_both", "src", "tgt", "mean", "both"),
    ("same_min_both", "tgt", "tgt", "minimum", "both"),
    ("same_centroid_both", "tgt", "tgt", "centroid", "both"),
]


def main() -> None:
    _, _, AD = ref_shim.load()
    dm = importlib.import_module("spatialcore.spatial.distance")
    coords, src, tgt = inputs.domains()
    out = {}
    for tag, sc, tc, metric, mode in CASES:
        obs = pd.DataFrame({"src": pd.Series(src, dtype=object), "tgt": pd.Series(tgt, dtype=object)})
        a = AD(np.zeros((coords.shape[0], 1), np.float32), obs=obs, obsm={"spatial": coords})
        dm.calculate_domain_distances(a, sc, tc, distance_metric=metric, output_mode=mode)
        if mode != "matrix":
            out[f"{tag}_dist"] = a.obs["distance_to_target"].to_numpy(dtype=np.float64)
            out[f"{tag}_nearest"] = np.array(["" if v is None or v != v else str(v) for v in a.obs["nearest_target_domain"]])
        M = dm.get_distance_matrix(a)
        out[f"{tag}_matrix"] = M.to_numpy(dtype=np.float64)
        out[f"{tag}_rows"] = np.array(list(M.index))
        out[f"{tag}_cols"] = np.array(list(M.columns))
    np.savez_compressed(os.path.join(HERE, "ref_distances.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
