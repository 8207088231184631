"""TEST INFRASTRUCTURE ONLY — CPU oracle for the spatial-statistics hot path.

Nothing under ``spatialcore_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, as the checker or as the CPU baseline — never as the product path.

Contents
--------
``restate.py``   numpy/scipy restatement of the reference algorithms, each function
                 citing the reference file:line it follows.
``moran_port.c`` C + OpenMP port of the Moran permutation loop (scanpy's numba kernel
                 as driven by squidpy), used for CPU timing on the GPU box.
``ref_shim.py``  loads the UNMODIFIED reference from /root/reference (this container
                 only) to pin the restatement and to generate ``tests/golden/*.npz``.

Parity status
-------------
* rows a1, a7-a12 (reference-own code) and the "next" rows ``identify_niches`` /
  ``calculate_domain_distances``: PINNED against outputs of the reference itself
  (``tests/golden/make_golden*.py`` -> ``tests/golden/ref_*.npz``).
* rows a2-a6 (the squidpy/scanpy segment ``morans_i`` delegates to): **parity unpinned**.
  squidpy (>=1.3.0) / scanpy (>=1.9.0) [R pyproject.toml:38-39] are not vendored in
  /root/reference and not installable here; the reference ships no tests or golden
  vectors.  The restatement follows their published algorithm
  and is cross-checked for internal consistency with the reference's own
  ``local_morans_i`` / ``build_spatial_weights`` outputs, and against known-answer vectors
  (closed-form Moran's I, analytic moments and z-scores on a 4-regular torus;
  ``tests/test_oracle.py::test_moran_known_answers_on_a_torus``).
"""
