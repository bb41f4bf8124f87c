"""CPU oracle for the BBMM exact-GP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-on-CPU, dtype generic, fp64 by
default in the tests) of the algorithms on the reference's hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
``gpytorch_b200`` never does (its ops raise when the CUDA library is absent).

Pinning status
--------------
* ``oracle.kernels`` (sq_dist / dist / RBF / Matern / scale) is PINNED: it is
  checked against outputs of the reference's own code executed in the build
  container (``tests/golden/make_golden.py`` runs
  ``gpytorch/functions/rbf_covariance.py``, ``matern_covariance.py`` and the
  ``sq_dist``/``dist`` bodies of ``gpytorch/kernels/kernel.py:26-60`` and stores
  the results in ``tests/golden/*.npz``) and against the known-answer matrices
  in ``test/kernels/test_rbf_kernel.py:105-142`` and
  ``test/kernels/test_matern_kernel.py:41-109``.
* ``oracle.linalg`` (mBCG ``linear_cg``, ``pivoted_cholesky``, the
  QR/Woodbury preconditioner, SLQ, ``inv_quad_logdet``, ``lanczos_tridiag``)
  restates the published algorithms of the third-party dependency
  ``linear_operator>=0.6.1`` (``setup.py:44``), whose source is NOT under
  ``/root/reference`` and which cannot be installed offline.  There is no
  reference golden vector for these pieces: **parity unpinned**.  They are
  anchored instead on dense-Cholesky ground truth (solve / log-det / MLL) and
  on the reference's call sites (``distributions/multivariate_normal.py:248-251``,
  ``variational/ciq_variational_strategy.py:56-64``).
* ``oracle/ski.py`` (SKI / KISS-GP, groundwork for SURVEY.md section 8f row 3): the interpolation and grid helpers are
  PINNED against outputs of the reference's own code (``tests/golden/ski_golden.npz``); the Toeplitz / Kronecker
  products restate linear_operator and are **parity unpinned** (checked against dense matrices).
"""
from . import kernels, linalg, mll, ski  # noqa: F401
