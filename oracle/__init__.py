"""
oracle -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (NumPy/SciPy + one small C file) of the reference's GP log-likelihood hot path
(ameli/gaussian-process-param-estimation, package ``gaussian_proc``), used as the checker for the CUDA product.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
it; nothing under ``gaussian-process-param-estimation_b200/`` does.

Parity status (see DESIGN.md "Oracle"):
* generate_correlation (dense + sparse), MixedCorrelation(eigenvalue|cholesky), DirectLikelihood, ProfileLikelihood,
  root finding: PINNED -- checked in tests/test_oracle_*.py against (a) the reference's shipped golden pickles
  (data/OptimalCovariance_With{,out}Prior.pickle, data/NoiseLevelResults.pickle; committed subsets in
  tests/golden/), (b) the compiled reference Cython generators in oracle/_ref (built by oracle/build_ref.py) and
  (c) vectors produced by importing the reference's own Python modules in the build container
  (tests/golden/make_golden.py).
* d/d(correlation_scale): the reference has no such derivative. parity unpinned by the reference;
  pinned here to Richardson finite differences of the pinned log-likelihood.
* hutchinson / slq estimators: parity unpinned (imate is an absent, unpinned third-party dependency and the
  reference's slq branches are dead code); the oracle supplies exact dense/sparse-LU values for the confidence band.
"""
