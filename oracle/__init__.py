"""CPU oracle for the BOSIP.jl hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / the reported CPU baseline.  The product
path (``bosip.jl_b200``) never imports it and fails loudly when its CUDA library is missing.

What it is: a float64 NumPy/SciPy restatement (``bosip_oracle.py``) of the reference's algorithm for
the path named in BASELINE.json, each function citing the reference file:line it follows, plus an
mpmath restatement at 50 digits (``mp_oracle.py``) used to generate ``tests/golden/*.json``.

Pinning status:
  * Likelihood algebra (loglike_marginal, log_marginal_likelihood_mean, log_likelihood_mean,
    log_approx_likelihood): PINNED against the reference's own fixtures
    (/root/reference/test/unit/test/posterior/likelihood.jl:1-136,
     /root/reference/test/unit/test/likelihoods/normal.jl:1-43) -- see tests/test_oracle_golden.py.
  * GP predict (BOSS.jl 0.6.x / AbstractGPs / KernelFunctions -- un-vendored registry dependencies,
    compat "0.6.1", no Manifest), log_sq_likelihood_mean, log_likelihood_variance, _log_prior,
    evidence, MaxVar argmax, calc_mmd, calc_tv: the reference holds no golden vector for them and
    cannot be executed in this image (no julia, no network), so for these rows the oracle restates
    the published formulas and is cross-checked only against the independent
    50-digit mpmath restatement: PARITY UNPINNED for those rows.
"""
