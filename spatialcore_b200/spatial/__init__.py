"""B200 drop-in for the hot path of ``spatialcore.spatial`` [R src/spatialcore/spatial/__init__.py:11-52].

In scope: neighbour graphs, global/local Moran's I, global/local Lee's L,
neighbourhood composition and the "next" rows of §8f: ``identify_niches`` (k-means on the profile
matrix) and ``calculate_domain_distances`` / ``get_distance_matrix`` (cross-set nearest neighbour and
pairwise distance reductions).  ``make_spatial_domains`` and ``get_domain_summary`` (R-side
computational geometry) are out of scope for this build.
"""

from spatialcore_b200.spatial.autocorrelation import (
    build_spatial_weights,
    lees_l,
    lees_l_local,
    lees_l_matrix,
    local_morans_i,
    morans_i,
    spatial_neighbors,
)
from spatialcore_b200.spatial.distance import calculate_domain_distances, get_distance_matrix
from spatialcore_b200.spatial.neighborhoods import compute_neighborhood_profile
from spatialcore_b200.spatial.niches import identify_niches

__all__ = [
    "morans_i",
    "local_morans_i",
    "lees_l",
    "lees_l_local",
    "lees_l_matrix",
    "build_spatial_weights",
    "spatial_neighbors",
    "compute_neighborhood_profile",
    "identify_niches",
    "calculate_domain_distances",
    "get_distance_matrix",
]
