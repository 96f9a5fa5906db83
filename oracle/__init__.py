"""CPU oracle for the nerf-dbr render/train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline) -- never as the thing shipped.  The product path
(``nerf_dbr_b200``) raises when its CUDA library is missing; it never routes
through this package.

Two restatements live here:

* ``oracle.nerf_oracle`` -- a torch-CPU restatement of the reference's render
  and train path (the reference's arithmetic *is* "torch CPU ops", so the
  faithful restatement keeps the same primitive ops and chunking).
* ``oracle/scalar_oracle.c`` -- a scalar C restatement of the pieces whose
  results are bit-exact targets (ray generation, sample placement, stratified
  jitter, inverse-CDF sampling, compositing with double-precision running
  products) plus a naive fp32 MLP for small cases; loaded through
  ``oracle.scalar``.

Pinning: the reference's own tests hold no golden vectors (shape/range checks
only), so parity is pinned by *running the reference itself* in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference`` and
writes the fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks both restatements against them.  The one exception is
``importance_sample`` (reference ``src/utils/rendering.py:54-100``), which is
dead code that raises at ``rendering.py:89``: its fixture comes from the
reference source with the documented one-line shape fix applied at generation
time, so that row is "parity pinned to reference + fix", not to the reference
as shipped.
"""
