"""CPU: the closed-form expectations the reference codes per likelihood (E[l], E[l^2], V[l] under y_j ~ N(mu_j, var_j) independent;
src/likelihoods/{normal,lognormal,normal_diff,normal_sum,exp}.jl) checked on the oracle against adaptive quadrature of the
plug-in likelihood itself.  The reference's tests never pin E[l^2] / V[l] numerically; this does, from
first principles."""
import math

import numpy as np
import pytest
from scipy.integrate import quad

from oracle import bosip_oracle as orc


def expect(f, mu, var):
    """E f(y), y ~ N(mu, var), by adaptive quadrature (the integrand is split at the mean and at +-2, 4, 8 standard deviations)."""
    sd = math.sqrt(var)
    g = lambda y: float(f(y)) * math.exp(-0.5 * ((y - mu) / sd) ** 2) / (sd * math.sqrt(2.0 * math.pi))
    pts = sorted(mu + k * sd for k in (-8, -4, -2, -1, 0, 1, 2, 4, 8))
    val, err = quad(g, mu - 14 * sd, mu + 14 * sd, points=pts, epsabs=0.0, epsrel=1e-13, limit=400)
    return val


def npdf(z, m, s):
    return np.exp(-0.5 * ((z - m) / s) ** 2) / (s * math.sqrt(2.0 * math.pi))


def check(terms, e1, e2):
    assert terms["log_like_mean"] == pytest.approx(math.log(e1), rel=1e-9, abs=1e-9)
    assert terms["log_sq_like_mean"] == pytest.approx(math.log(e2), rel=1e-9, abs=1e-9)
    v = e2 - e1 * e1
    if v > 1e-6 * e2:
        assert terms["log_like_var"] == pytest.approx(math.log(v), rel=1e-7, abs=1e-7)


@pytest.mark.parametrize("seed", range(6))
def test_normal_family_expectations(seed):
    rng = np.random.default_rng(seed)
    p = int(rng.integers(1, 4))
    mu = rng.normal(0, 1.5, p); var = rng.uniform(0.05, 2.0, p); z = rng.normal(0, 1.5, p); s0 = rng.uniform(0.4, 2.0, p)
    # NormalLikelihood: l(y) = prod_j N(z_j; y_j, s0_j)                                                   normal.jl:28-79
    e1 = math.prod(expect(lambda y, j=j: npdf(z[j], y, s0[j]), mu[j], var[j]) for j in range(p))
    e2 = math.prod(expect(lambda y, j=j: npdf(z[j], y, s0[j]) ** 2, mu[j], var[j]) for j in range(p))
    check(orc.like_terms(orc.NormalLikelihood(z, s0), mu, var), e1, e2)
    t = orc.like_terms(orc.NormalLikelihood(z, s0), mu, var)
    assert orc.log_likelihood_mean(orc.NormalLikelihood(z, s0), mu[:, None], np.sqrt(var)[:, None])[0] == pytest.approx(t["log_like_mean"], rel=1e-12)
    assert orc.log_sq_likelihood_mean(orc.NormalLikelihood(z, s0), mu[:, None], np.sqrt(var)[:, None])[0] == pytest.approx(t["log_sq_like_mean"], rel=1e-12)
    # NormalDiffLikelihood: the same with z = 0                                                            normal_diff.jl:17-73
    e1 = math.prod(expect(lambda y, j=j: npdf(0.0, y, s0[j]), mu[j], var[j]) for j in range(p))
    e2 = math.prod(expect(lambda y, j=j: npdf(0.0, y, s0[j]) ** 2, mu[j], var[j]) for j in range(p))
    check(orc.like_terms(orc.NormalDiffLikelihood(s0), mu, var), e1, e2)
    # LogNormalLikelihood: l(y) = prod_j LogNormal(y_j - s_j^2/2, s_j).pdf(z_j), s_j^2 = log(1 + CV_j^2)      lognormal.jl:26-91
    cv = rng.uniform(0.3, 1.2, p); zpos = np.exp(rng.normal(0, 0.7, p)); sl = np.sqrt(np.log(1 + cv ** 2))
    f = lambda y, j: npdf(math.log(zpos[j]), y - sl[j] ** 2 / 2, sl[j]) / zpos[j]
    e1 = math.prod(expect(lambda y, j=j: f(y, j), mu[j], var[j]) for j in range(p))
    e2 = math.prod(expect(lambda y, j=j: f(y, j) ** 2, mu[j], var[j]) for j in range(p))
    check(orc.like_terms(orc.LogNormalLikelihood(np.log(zpos), cv), mu, var), e1, e2)


@pytest.mark.parametrize("seed", range(4))
def test_normal_sum_and_exp_expectations(seed):
    rng = np.random.default_rng(100 + seed)
    lengths = [int(k) for k in rng.integers(1, 4, size=int(rng.integers(1, 3)))]
    P = sum(lengths)
    mu = rng.normal(0, 1.0, P); var = rng.uniform(0.05, 1.0, P)
    z = rng.normal(0, 1.5, len(lengths)); s0 = rng.uniform(0.5, 2.0, len(lengths))
    # NormalSumLikelihood: l = prod_g N(z_g; sum of the group's outputs, s0_g); the sum of independent normals is normal   normal_sum.jl:18-115
    ms = np.array([mu[sum(lengths[:g]):sum(lengths[:g + 1])].sum() for g in range(len(lengths))])
    vs = np.array([var[sum(lengths[:g]):sum(lengths[:g + 1])].sum() for g in range(len(lengths))])
    e1 = math.prod(expect(lambda y, g=g: npdf(z[g], y, s0[g]), ms[g], vs[g]) for g in range(len(lengths)))
    e2 = math.prod(expect(lambda y, g=g: npdf(z[g], y, s0[g]) ** 2, ms[g], vs[g]) for g in range(len(lengths)))
    check(orc.like_terms(orc.NormalSumLikelihood(lengths, z, s0), mu, var), e1, e2)
    # ExpLikelihood: l(y) = exp(y), one output                                                                exp.jl:7-63
    m1, v1 = float(rng.normal(-1, 1)), float(rng.uniform(0.05, 1.5))
    check(orc.like_terms(orc.ExpLikelihood(), np.array([m1]), np.array([v1])), expect(np.exp, m1, v1), expect(lambda y: np.exp(2 * y), m1, v1))


def test_prior_logpdfs_against_scipy():
    """_log_prior (src/posterior/log_posterior.jl:228-229) over products of Uniform / Normal / truncated(Normal): the oracle's
    closed forms against scipy.stats (an independent implementation of the same Distributions.jl semantics: -Inf outside the
    support, boundaries included, truncation renormalised)."""
    from scipy import stats
    rng = np.random.default_rng(3)
    kinds = [orc.UNIFORM, orc.NORMAL, orc.TRUNCNORMAL, orc.TRUNCNORMAL]
    prm = [(-5.0, 5.0), (0.3, 1.7), (0.0, 5.0 / 3.0, -5.0, 5.0), (4.0, 0.5, -5.0, 5.0)]      # the last one is truncated well inside its bulk
    prior = orc.ProductPrior(kinds, prm)
    X = rng.uniform(-6, 6, (4, 400))
    X[:, 0] = [-5.0, 0.0, 5.0, -5.0]                                                        # support boundaries are inside
    ref = (stats.uniform(-5.0, 10.0).logpdf(X[0]) + stats.norm(0.3, 1.7).logpdf(X[1])
           + stats.truncnorm((-5.0 - 0.0) / (5.0 / 3.0), (5.0 - 0.0) / (5.0 / 3.0), 0.0, 5.0 / 3.0).logpdf(X[2])
           + stats.truncnorm((-5.0 - 4.0) / 0.5, (5.0 - 4.0) / 0.5, 4.0, 0.5).logpdf(X[3]))
    got = orc.log_prior(prior, X)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and fin[0] and (~fin).sum() > 50
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-12, atol=1e-12)
    assert orc.log_prior(prior, X[:, 0]) == pytest.approx(ref[0], rel=1e-12)                 # vector method


@pytest.mark.parametrize("kernel_id", [orc.SE, orc.MATERN32, orc.MATERN52])
def test_gp_predict_against_sklearn(kernel_id):
    """BOSS.mean / var (AbstractGPs posterior of alpha^2 k(./lambda) with noise sigma^2 on the training diagonal)
    against scikit-learn's GaussianProcessRegressor with the hyper-parameters fixed -- an independent implementation of the same
    textbook formulas and of the ARD SE / Matern-3/2 / Matern-5/2 kernels."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    rng = np.random.default_rng(11 + kernel_id)
    d, N, M = 3, 40, 25
    X = rng.uniform(-5, 5, (d, N)); Y = (np.sin(X).sum(0) + 0.1 * rng.standard_normal(N))[None, :]
    lam = rng.uniform(1.0, 4.0, (d, 1)); alpha = np.array([1.3]); sigma = np.array([0.2])
    Xq = rng.uniform(-5, 5, (d, M)); Xq[:, :3] = X[:, :3] + 1e-3
    mu, var = orc.gp_predict(orc.gp_fit(kernel_id, X, Y, lam, alpha, sigma), Xq)
    base = RBF(length_scale=lam[:, 0]) if kernel_id == orc.SE else Matern(length_scale=lam[:, 0], nu=1.5 if kernel_id == orc.MATERN32 else 2.5)
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(alpha[0] ** 2, "fixed") * base, alpha=sigma[0] ** 2, optimizer=None).fit(X.T, Y[0])
    m_sk, s_sk = gpr.predict(Xq.T, return_std=True)
    np.testing.assert_allclose(mu[0], m_sk, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(var[0], s_sk ** 2, rtol=1e-7, atol=1e-10 * alpha[0] ** 2)


def test_mmd_against_gaussian_closed_form():
    """calc_mmd (src/metrics/mmd.jl:23-28) on samples of two isotropic Gaussians against the population value: for the SE kernel
    E k(x, y) = prod_k lam_k / sqrt(lam_k^2 + s1^2 + s2^2) * exp(-(m1 - m2)^2 / (2 (lam_k^2 + s1^2 + s2^2))).  The V-statistic
    includes the n + m unit diagonal terms, which the expectation below accounts for."""
    rng = np.random.default_rng(0)
    d, n, m = 2, 12000, 10000
    lam = np.array([1.0, 1.5]); m1, s1, m2, s2 = 0.0, 1.0, 0.4, 1.2
    A = m1 + s1 * rng.standard_normal((d, n)); B = m2 + s2 * rng.standard_normal((d, m))
    ek = lambda dm, sa, sb: float(np.prod(lam / np.sqrt(lam ** 2 + sa ** 2 + sb ** 2) * np.exp(-dm ** 2 / (2 * (lam ** 2 + sa ** 2 + sb ** 2)))))
    exx, eyy, exy = ek(0.0, s1, s1), ek(0.0, s2, s2), ek(m1 - m2, s1, s2)
    expected = (exx * (n - 1) / n + 1.0 / n) + (eyy * (m - 1) / m + 1.0 / m) - 2.0 * exy
    got = orc.calc_mmd(orc.SE, lam, A, B)[0]
    assert got == pytest.approx(expected, abs=6e-3) and expected > 0.03          # sampling noise of the estimate is ~2e-3 here
