"""CPU oracle for the mfBiclust factorization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.

The reference (jonalim/mfBiclust) is pure R and R is not installed in this
image, and its arithmetic lives in un-vendored CRAN/Bioconductor packages
(NMF 0.21.0 ``.fcnnls``, nipals 0.5, MASS 7.3-51.1 ``ginv``, EBImage 4.24.0
``otsu``, base R 3.6.0 ``prcomp``/``hist``/``max.col``/``runif``/``sample``;
versions from ``packrat/packrat.lock``).  Every module here is a NumPy
restatement of the published algorithm, anchored on the reference's own call
sites (cited ``file:line`` relative to the reference root) and pinned against
the reference's artefacts: ``tests/testdata/ref_bces.rda`` and the rendered
vignette ``vignettes/my-vignette.html`` (see ``tests/golden/`` and
``tests/test_oracle_*.py``).  Paths with no artefact say "parity unpinned" in
their docstring.
"""
