"""GPU: the reference's own unit tests for this path, re-run through the C ABI.

Mirrors /root/reference/test/unit/test/posterior/likelihood.jl:1-136 and test/unit/test/likelihoods/normal.jl:1-43 with
the same MockModelPosterior (test/unit/utils.jl:1-11), the same fixture and the same assertions; expected values come
from the oracle and from tests/golden/likelihood_fixture.json.  `≈` in the reference is isapprox with rtol=sqrt(eps);
here the bar is 1e-10 relative."""
import json
import math
import os

import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-10
FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "likelihood_fixture.json")))


class MockModelPosterior:
    """test/unit/utils.jl:3-11"""

    def __init__(self, mean_val, var_val):
        self.mean_val = np.asarray(mean_val, float); self.var_val = np.asarray(var_val, float)

    def mean(self, X):
        return self.mean_val if np.ndim(X) == 1 else np.repeat(self.mean_val[:, None], np.shape(X)[1], 1)

    def var(self, X):
        return self.var_val if np.ndim(X) == 1 else np.repeat(self.var_val[:, None], np.shape(X)[1], 1)


z_obs, std_obs = np.array(FX["z_obs"]), np.array(FX["std_obs"])
mean_val = np.array(FX["mean"])
var_zero, var_nonzero = np.zeros(3), std_obs ** 2
x = np.array([0.0]); X = np.array([[0.0, 1.0, 2.0]])
obs_dim, n_points = 3, 3
olike = orc.NormalLikelihood(z_obs, std_obs)
expected_lm = lambda yi: orc.normlogpdf(yi, std_obs, z_obs)
expected_l = lambda yi: expected_lm(yi).sum()
expected_ml_mean = lambda mu, s: orc.normlogpdf(mu, np.sqrt(std_obs ** 2 + s ** 2), z_obs)
expected_l_mean = lambda mu, s: expected_ml_mean(mu, s).sum()


@pytest.fixture(scope="module")
def env(bb, ctx):
    bb.api._default_ctx = ctx
    like = bb.NormalLikelihood(z_obs, std_obs)
    return bb, like, MockModelPosterior(mean_val, var_zero), MockModelPosterior(mean_val, var_nonzero)


def test_log_approx_marginal_likelihood(env):  # likelihood.jl:24-40
    bb, like, post_zero, _ = env
    f = bb.log_approx_marginal_likelihood(like, post_zero)
    assert callable(f)
    r = f(x)
    assert r.shape == (obs_dim,)
    np.testing.assert_allclose(r, expected_lm(mean_val), rtol=RTOL)
    np.testing.assert_allclose(r, FX["cases"]["var_zero"]["loglike_marginal"], rtol=RTOL)
    rm = f(X)
    assert rm.shape == (obs_dim, n_points)
    for i in range(n_points):
        np.testing.assert_allclose(rm[:, i], f(X[:, i]), rtol=RTOL)


def test_log_approx_likelihood(env):  # likelihood.jl:42-57
    bb, like, post_zero, _ = env
    f = bb.log_approx_likelihood(like, post_zero)
    r = f(x)
    assert isinstance(r, float) and r == pytest.approx(expected_l(mean_val), rel=RTOL)
    rv = f(X)
    assert rv.shape == (n_points,)
    for i in range(n_points):
        assert rv[i] == pytest.approx(f(X[:, i]), rel=RTOL)


def test_marginal_sum_consistency(env):  # likelihood.jl:59-65, 120-126
    bb, like, post_zero, post_nonzero = env
    f_ml, f_l = bb.log_approx_marginal_likelihood(like, post_zero), bb.log_approx_likelihood(like, post_zero)
    assert f_ml(x).sum() == pytest.approx(f_l(x), rel=RTOL)
    np.testing.assert_allclose(f_ml(X).sum(0), f_l(X), rtol=RTOL)
    g_ml, g_l = bb.log_marginal_likelihood_mean(like, post_nonzero), bb.log_likelihood_mean(like, post_nonzero)
    assert g_ml(x).sum() == pytest.approx(g_l(x), rel=RTOL)
    np.testing.assert_allclose(g_ml(X).sum(0), g_l(X), rtol=RTOL)


def test_log_marginal_likelihood_mean(env):  # likelihood.jl:67-94
    bb, like, post_zero, post_nonzero = env
    f_zero, f_nonzero = bb.log_marginal_likelihood_mean(like, post_zero), bb.log_marginal_likelihood_mean(like, post_nonzero)
    rz = f_zero(x)
    assert rz.shape == (obs_dim,)
    np.testing.assert_allclose(rz, expected_lm(mean_val), rtol=RTOL)
    rn = f_nonzero(x)
    np.testing.assert_allclose(rn, expected_ml_mean(mean_val, std_obs), rtol=RTOL)
    np.testing.assert_allclose(rn, FX["cases"]["var_nonzero"]["log_ml_mean"], rtol=RTOL)
    assert not np.allclose(rn, rz)
    rm = f_zero(X)
    assert rm.shape == (obs_dim, n_points)
    for i in range(n_points):
        np.testing.assert_allclose(rm[:, i], f_zero(X[:, i]), rtol=RTOL)


def test_log_likelihood_mean(env):  # likelihood.jl:96-118
    bb, like, post_zero, post_nonzero = env
    f_zero, f_nonzero = bb.log_likelihood_mean(like, post_zero), bb.log_likelihood_mean(like, post_nonzero)
    assert f_zero(x) == pytest.approx(expected_l_mean(mean_val, np.zeros(3)), rel=RTOL)
    assert f_nonzero(x) == pytest.approx(expected_l_mean(mean_val, std_obs), rel=RTOL)
    assert f_nonzero(x) == pytest.approx(FX["cases"]["var_nonzero"]["log_like_mean"], rel=RTOL)
    assert f_nonzero(x) != pytest.approx(f_zero(x))
    rv = f_zero(X)
    assert rv.shape == (n_points,)
    for i in range(n_points):
        assert rv[i] == pytest.approx(f_zero(X[:, i]), rel=RTOL)


def test_zero_variance_approx_equals_mean(env):  # likelihood.jl:128-136
    bb, like, post_zero, _ = env
    np.testing.assert_allclose(bb.log_approx_marginal_likelihood(like, post_zero)(x), bb.log_marginal_likelihood_mean(like, post_zero)(x), rtol=RTOL)
    assert bb.log_approx_likelihood(like, post_zero)(x) == pytest.approx(bb.log_likelihood_mean(like, post_zero)(x), rel=RTOL)


def test_sq_mean_and_variance_pins(env):  # SURVEY.md Appendix C rows 4-5 (never pinned numerically by the reference)
    bb, like, post_zero, post_nonzero = env
    c = FX["cases"]["var_nonzero"]
    assert bb.log_sq_likelihood_mean(like, post_nonzero)(x) == pytest.approx(c["log_sq_like_mean"], rel=RTOL)
    assert bb.log_likelihood_variance(like, post_nonzero)(x) == pytest.approx(c["log_like_var"], rel=RTOL)
    rv = bb.log_likelihood_variance(like, post_nonzero)(X)
    assert rv.shape == (n_points,) and np.allclose(rv, c["log_like_var"], rtol=RTOL)
    assert bb.log_sq_likelihood_mean(like, post_zero)(x) == pytest.approx(FX["cases"]["var_zero"]["log_sq_like_mean"], rel=RTOL)
    lv0 = bb.log_likelihood_variance(like, post_zero)(x)   # exp(a) - exp(2b) with a == 2b up to rounding: 0 or a tiny residue
    assert lv0 == -np.inf or lv0 < -30


def test_normal_jl(env):  # test/unit/test/likelihoods/normal.jl:11-43
    bb, like, _, _ = env
    y = np.array([1.1, 2.2, 3.3]); Y = np.array([[1.1, 0.5], [2.2, 1.5], [3.3, 2.5]])
    r = bb.loglike_marginal(like, y)
    assert r.shape == (3,)
    np.testing.assert_allclose(r, expected_lm(y), rtol=RTOL)
    rm = bb.loglike_marginal(like, Y)
    assert rm.shape == (3, 2)
    np.testing.assert_allclose(rm[:, 0], bb.loglike_marginal(like, Y[:, 0]), rtol=RTOL)
    np.testing.assert_allclose(rm[:, 1], bb.loglike_marginal(like, Y[:, 1]), rtol=RTOL)
    s = bb.loglike(like, y)
    assert isinstance(s, float) and s == pytest.approx(expected_l(y), rel=RTOL)
    sv = bb.loglike(like, Y)
    assert sv.shape == (2,)
    assert sv[0] == pytest.approx(bb.loglike(like, Y[:, 0]), rel=RTOL) and sv[1] == pytest.approx(bb.loglike(like, Y[:, 1]), rel=RTOL)
    assert bb.loglike_marginal(like, y).sum() == pytest.approx(s, rel=RTOL)
    np.testing.assert_allclose(bb.loglike_marginal(like, Y).sum(0), sv, rtol=RTOL)


def test_likelihood_layer_against_oracle_random(env):
    bb, _, _, _ = env
    rng = np.random.default_rng(5)
    p, M = 4, 5000
    like = bb.NormalLikelihood(rng.normal(0, 1, p), rng.uniform(0.1, 2, p)); ol = orc.NormalLikelihood(like.z_obs, like.std_obs)
    mu = rng.normal(0, 1.5, (p, M)); var = rng.uniform(0, 3, (p, M)); var[:, :100] = 0.0
    r = bb.normal_likelihood_eval(like, mu, var)
    std = np.sqrt(var)
    np.testing.assert_allclose(r["loglike_marginal"], orc.loglike_marginal(ol, mu), rtol=RTOL)
    np.testing.assert_allclose(r["loglike"], orc.loglike(ol, mu), rtol=RTOL)
    np.testing.assert_allclose(r["log_marginal_mean"], orc.log_marginal_likelihood_mean(ol, mu, std), rtol=RTOL)
    np.testing.assert_allclose(r["log_like_mean"], orc.log_likelihood_mean(ol, mu, std), rtol=RTOL)
    np.testing.assert_allclose(r["log_sq_like_mean"], orc.log_sq_likelihood_mean(ol, mu, std), rtol=RTOL)
    ov = orc.log_likelihood_variance(ol, mu, std)
    fin = np.isfinite(ov) & (var.min(0) > 1e-3)
    np.testing.assert_allclose(r["log_like_var"][fin], ov[fin], rtol=1e-9)
    assert np.all(np.isneginf(r["log_like_var"][:100]) | (r["log_like_var"][:100] < -25))


def test_domain_error_like_the_reference(env):  # src/posterior/likelihood.jl:61-65
    """A NEGATIVE predictive variance (which a non-clamping ModelPosterior may return) makes E[l^2] < E[l]^2 by more than
    1e-8 -> DomainError in the reference; here BOSIP_B200_EDOMAIN -> DomainError."""
    bb, _, _, _ = env
    like = bb.NormalLikelihood([0.0], [1.0])
    mu, var = np.array([[0.0, 1.0]]), np.array([[0.5, -0.4]])
    # the same V in the oracle's statement of _assure_nonneg: exp(log E[l^2]) - exp(2 log E[l]) with var = -0.4
    so2, v = 1.0, -0.4
    log_sq = -math.log(2 * math.sqrt(math.pi)) + float(orc.normlogpdf(1.0, math.sqrt((so2 + 2 * v) / 2), 0.0))
    log_L = float(orc.normlogpdf(1.0, math.sqrt(so2 + v), 0.0))
    with pytest.raises(orc.DomainError):
        orc.assure_nonneg(np.array([math.exp(log_sq) - math.exp(2 * log_L)]))
    with pytest.raises(bb.DomainError) as e:
        bb.normal_likelihood_eval(like, mu, var, ("log_like_var",))
    assert e.value.code == bb._lib.EDOMAIN and "index 1" in str(e.value)
    # tiny negative values clamp to zero -> -Inf, counted
    r = bb.normal_likelihood_eval(bb.NormalLikelihood([0.0], [1.0]), np.array([[30.0]]), np.array([[-1e-12]]), ("log_like_var",))
    assert r["log_like_var"][0] == -np.inf


# ======================================================================================================================
# The other closed-form likelihoods: test/unit/test/likelihoods/{lognormal,normal_diff,
# normal_sum,exp}.jl re-run through the C ABI with the same fixtures + random parity against the oracle restatements.
# ======================================================================================================================
def _family_cases(bb):
    rng = np.random.default_rng(11)
    return [
        ("lognormal", bb.LogNormalLikelihood([0.0, 0.5, 1.0], [0.1, 0.2, 0.1]), orc.LogNormalLikelihood([0.0, 0.5, 1.0], [0.1, 0.2, 0.1]),
         np.array([0.1, 0.6, 1.1])),                                                         # lognormal.jl:1-13
        ("normal_diff", bb.NormalDiffLikelihood([0.5, 1.0, 2.0]), orc.NormalDiffLikelihood([0.5, 1.0, 2.0]), np.array([0.1, 0.2, 0.3])),
        ("normal_sum", bb.NormalSumLikelihood([2, 1, 3], [1.0, 2.0, 3.0], [0.5, 1.0, 2.0]), orc.NormalSumLikelihood([2, 1, 3], [1.0, 2.0, 3.0], [0.5, 1.0, 2.0]),
         rng.normal(0.5, 0.3, 6)),
        ("exp", bb.ExpLikelihood(), orc.ExpLikelihood(), np.array([2.5])),                   # exp.jl:1-4
    ]


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_likelihood_family_fixture_tests(env, idx):
    bb = env[0]
    name, like, olike, y = _family_cases(bb)[idx]
    p = len(y)
    var = np.full(p, 0.25) if name != "lognormal" else olike.sigma_log ** 2          # lognormal.jl:52-53, exp.jl:17-19
    post_zero, post_nonzero = MockModelPosterior(y, np.zeros(p)), MockModelPosterior(y, var)
    X3 = np.array([[0.0, 1.0, 2.0]])
    # loglike_marginal / loglike, vector and matrix methods + consistency
    Y = np.stack([y, y - 0.2], axis=1)
    lm, l = bb.loglike_marginal(like, y), bb.loglike(like, y)
    ot = orc.like_terms(olike, y, None)
    np.testing.assert_allclose(lm, ot["loglike_marginal"], rtol=RTOL)
    assert isinstance(l, float) and l == pytest.approx(float(ot["loglike"]), rel=RTOL)
    Lm = bb.loglike_marginal(like, Y)
    assert Lm.shape == (like.obs_dim, 2)
    np.testing.assert_allclose(Lm[:, 0], lm, rtol=RTOL)
    np.testing.assert_allclose(Lm.sum(0), bb.loglike(like, Y), rtol=RTOL)
    assert lm.sum() == pytest.approx(l, rel=RTOL)
    # expectations: zero variance == plug-in, nonzero matches the closed form, matrix == columnwise, marginal/sum consistency
    f_zero, f_non = bb.log_marginal_likelihood_mean(like, post_zero), bb.log_marginal_likelihood_mean(like, post_nonzero)
    np.testing.assert_allclose(f_zero(x), lm, rtol=RTOL)
    on = orc.like_terms(olike, y, var)
    np.testing.assert_allclose(f_non(x), on["log_marginal_mean"], rtol=RTOL)
    assert not np.allclose(f_non(x), f_zero(x))
    Rm = f_non(X3)
    assert Rm.shape == (like.obs_dim, 3)
    for i in range(3):
        np.testing.assert_allclose(Rm[:, i], f_non(X3[:, i]), rtol=RTOL)
    g = bb.log_likelihood_mean(like, post_nonzero)
    assert g(x) == pytest.approx(float(on["log_like_mean"]), rel=RTOL)
    np.testing.assert_allclose(Rm.sum(0), g(X3), rtol=RTOL)
    # variance: type/shape/self-consistency as the reference checks it, plus the oracle's number
    h = bb.log_likelihood_variance(like, post_nonzero)
    assert isinstance(h(x), float)
    hv = h(X3)
    assert hv.shape == (3,)
    for i in range(3):
        assert hv[i] == pytest.approx(h(X3[:, i]), rel=RTOL)
    assert h(x) == pytest.approx(float(on["log_like_var"]), rel=1e-9)
    if name == "exp":                                                                   # exp.jl:22-46
        assert bb.log_likelihood_mean(like, post_zero)(x) == pytest.approx(2.5, rel=RTOL)
        assert g(x) == pytest.approx(2.5 + 0.5 * 0.25, rel=RTOL) and g(x) > bb.log_likelihood_mean(like, post_zero)(x)
        assert np.exp(h(x)) > 0


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_likelihood_family_random_parity_and_gp_path(env, ctx, idx):
    bb = env[0]
    name, like, olike, y = _family_cases(bb)[idx]
    rng = np.random.default_rng(40 + idx)
    p, M = like.model_dim, 3000
    mu = rng.normal(float(np.mean(y)), 0.4, (p, M)); var = rng.uniform(0.0, 0.5, (p, M)); var[:, :50] = 0.0
    r = bb.normal_likelihood_eval(like, mu, var)
    o = orc.like_terms(olike, mu, var)
    for k in ("loglike_marginal", "loglike", "log_marginal_mean", "log_like_mean", "log_sq_like_mean"):
        np.testing.assert_allclose(r[k], o[k], rtol=RTOL, atol=1e-13)
    fin = np.isfinite(o["log_like_var"]) & (var.min(0) > 1e-3)
    np.testing.assert_allclose(r["log_like_var"][fin], o["log_like_var"][fin], rtol=1e-8)
    # the fused GP path with this likelihood == GP predict (oracle) + the likelihood restatement + prior
    pr = bb.synth.make_problem("X", d=2, p=p, N=60, seed=900 + idx)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, 0.3 * pr.alpha, 0.3 * pr.sigma)
    Ysc = 0.3 * pr.Y + float(np.mean(y))
    post = bb.model_posterior(model, pr.X, Ysc, ctx=ctx)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, Ysc, pr.lam, 0.3 * pr.alpha, 0.3 * pr.sigma)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params); oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(2, 2000))
    res = bb.posterior_eval(post, like, prior, Xq, 7)
    omu, ovar = orc.gp_predict(ofit, Xq)
    ot = orc.like_terms(olike, omu, ovar); lp = orc.log_prior(oprior, Xq)
    np.testing.assert_allclose(res["log_approx_post"], lp + ot["loglike"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(res["log_post_mean"], lp + ot["log_like_mean"], rtol=RTOL, atol=1e-12)
    ref = 2 * lp + ot["log_like_var"]
    fin = np.isfinite(ref) & np.isfinite(res["log_post_var"])
    assert fin.mean() > 0.9
    assert np.all(np.abs(res["log_post_var"][fin] - ref[fin]) <= 1e-10 * np.abs(ref[fin]) + 1e-6)
