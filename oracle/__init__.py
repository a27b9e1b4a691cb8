"""CPU oracle for the curvature hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-torch (CPU, any float dtype; tests use float64) restatement of the
arithmetic the reference performs on the path named by ``BASELINE.json``:

* ``curvature_oracle`` -- Jacobians, full/diag GGN, full/diag EF and the KFAC
  factor semantics of the curvlinops adapter
  (reference ``laplace/curvature/curvature.py``, ``laplace/curvature/curvlinops.py``).
* ``kron_oracle``      -- ``Kron.decompose`` / ``KronDecomposed`` algebra and the
  Full/Diag posterior predictive maths (reference ``laplace/utils/matrix.py``,
  ``laplace/utils/utils.py``, ``laplace/baselaplace.py``).

Pinning status
--------------
* Jacobians, full/diag GGN, full/diag EF, eigendecomposition, Kron predictive,
  Full/Diag predictive: **pinned** against the unmodified reference imported in the
  build container (``oracle/ref_shim.py``) and against the committed golden
  vectors in ``tests/golden/`` that were generated from it
  (``tests/golden/make_golden.py``).
* KFAC factor arithmetic: the reference delegates it to the un-vendored
  dependency ``curvlinops-for-pytorch==2.0.0`` (``uv.lock:321-322``), which is
  absent here.  The restatement follows that library's published KFAC-expand /
  KFAC-reduce definitions and is pinned only through the identities the
  reference's own tests use (``tests/test_curv_backends_curvlinops.py``):
  exactness of every diagonal GGN block for a single datum, exactness of bias
  blocks, batch additivity, data-count normalisation.  **parity unpinned** with
  respect to curvlinops itself.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline
legs may import anything from this package.
"""
