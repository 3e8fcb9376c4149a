"""CPU: identities the reference's own tests assert (marginal-sum consistency, zero variance => mean == plug-in,
matrix == column-wise; test/unit/test/posterior/likelihood.jl:59-65,120-136) and the size-independent properties of the
metrics, checked on the oracle over random inputs (hypothesis)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import bosip_oracle as orc

finite = st.floats(-5.0, 5.0, allow_nan=False, allow_infinity=False)
pos = st.floats(0.05, 3.0, allow_nan=False, allow_infinity=False)


@st.composite
def like_case(draw):
    p = draw(st.integers(1, 4)); M = draw(st.integers(1, 6))
    z = np.array(draw(st.lists(finite, min_size=p, max_size=p))); so = np.array(draw(st.lists(pos, min_size=p, max_size=p)))
    mu = np.array(draw(st.lists(finite, min_size=p * M, max_size=p * M))).reshape(p, M)
    sd = np.array(draw(st.lists(pos, min_size=p * M, max_size=p * M))).reshape(p, M)
    return orc.NormalLikelihood(z, so), mu, sd


@settings(max_examples=60, deadline=None, derandomize=True)
@given(like_case())
def test_likelihood_identities(c):
    like, mu, sd = c
    # sum of the marginals == joint (likelihood.jl:59-65, 120-126)
    np.testing.assert_allclose(orc.loglike_marginal(like, mu).sum(0), orc.loglike(like, mu), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.log_marginal_likelihood_mean(like, mu, sd).sum(0), orc.log_likelihood_mean(like, mu, sd), rtol=1e-12, atol=1e-12)
    # zero predictive variance: the expected likelihood is the plug-in likelihood (likelihood.jl:128-136)
    np.testing.assert_allclose(orc.log_likelihood_mean(like, mu, np.zeros_like(sd)), orc.loglike(like, mu), rtol=1e-12, atol=1e-12)
    # matrix method == column-wise vector method
    for i in range(mu.shape[1]):
        np.testing.assert_allclose(orc.log_likelihood_mean(like, mu[:, i:i + 1], sd[:, i:i + 1])[0], orc.log_likelihood_mean(like, mu, sd)[i], rtol=1e-13)
    # E[l^2] >= E[l]^2: the variance of the likelihood is never below the clamp window (src/posterior/likelihood.jl:61-65)
    lv = orc.log_likelihood_variance(like, mu, sd)
    assert np.all((lv == -np.inf) | np.isfinite(lv))
    assert np.all(orc.log_sq_likelihood_mean(like, mu, sd) >= 2.0 * orc.log_likelihood_mean(like, mu, sd) - 1e-9)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_tv_properties(G, seed):
    rng = np.random.default_rng(seed)
    lw = rng.normal(0, 0.3, G); a = rng.normal(-3, 2, G); t = rng.normal(-3, 2, G)
    tv = orc.calc_tv(lw, a, t)
    assert 0.0 <= tv <= 1.0 + 1e-12
    assert orc.calc_tv(lw, a, a) == 0.0
    assert tv == orc.calc_tv(lw, t, a)                                         # symmetric in the two densities
    np.testing.assert_allclose(orc.calc_tv(lw, a + 7.5, t - 3.25), tv, rtol=1e-10, atol=1e-15)   # both are normalised first (tv.jl:64-70)
    np.testing.assert_allclose(orc.calc_tv(lw + 2.0, a, t), tv, rtol=1e-10, atol=1e-15)           # a common weight factor cancels


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(1, 4), st.integers(1, 30), st.integers(1, 30), st.sampled_from([orc.SE, orc.MATERN32, orc.MATERN52]), st.integers(0, 2 ** 31 - 1))
def test_mmd_properties(d, n, m, kernel_id, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, n)); B = 0.3 + rng.standard_normal((d, m)); lam = rng.uniform(0.5, 2.0, d)
    v, sAA, sBB, sAB = orc.calc_mmd(kernel_id, lam, A, B)
    assert v >= -1e-12                                                          # biased V-statistic of a positive-definite kernel
    assert abs(orc.calc_mmd(kernel_id, lam, A, A)[0]) < 1e-13
    np.testing.assert_allclose(orc.calc_mmd(kernel_id, lam, B, A)[0], v, rtol=1e-10, atol=1e-14)
    shift = rng.uniform(-3, 3, (d, 1))
    np.testing.assert_allclose(orc.calc_mmd(kernel_id, lam, A + shift, B + shift)[0], v, rtol=1e-8, atol=1e-13)   # stationary kernels
    assert sAA >= n - 1e-9 and sBB >= m - 1e-9                                  # the diagonals (k = 1) are included (mmd.jl:23-28)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.integers(1, 50), st.integers(0, 2 ** 31 - 1))
def test_weighted_quantile_monotone(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=n); w = rng.uniform(0.1, 2.0, n)
    qs = [orc.weighted_quantile(v, w, p) for p in (0.0, 0.1, 0.5, 0.9, 1.0)]
    assert all(a <= b + 1e-15 for a, b in zip(qs, qs[1:]))
    assert abs(qs[0] - v.min()) < 1e-12 and abs(qs[-1] - v.max()) < 1e-12
