"""Float64 NumPy/SciPy restatement of the BOSIP.jl hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference) or, for the arithmetic that
lives in the un-vendored dependencies BOSS.jl 0.6.x / AbstractGPs / KernelFunctions 0.10.x /
Distributions 0.25.x / StatsFuns, the published formula it restates.

Array conventions follow the reference: inputs ``X`` are d x M (one column per point), GP outputs
are p x M (test/unit/utils.jl:8-11).  NumPy arrays are C-ordered, so a Julia column-major d x M
matrix is passed here as an array of shape (M, d) would be in memory; to keep the *indexing* of the
reference readable all functions below take/return arrays indexed exactly like Julia: ``X[k, i]`` is
coordinate k of point i, ``mu[j, i]`` is output j at point i.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.special import ndtr

SE, MATERN32, MATERN52 = 0, 1, 2
KERNEL_NAMES = {SE: "SqExponential", MATERN32: "Matern32", MATERN52: "Matern52"}
LOG2PI = math.log(2.0 * math.pi)
MAX_NEG_VAR = 1e-8  # src/posterior/likelihood.jl:3


class DomainError(ValueError):
    """Mirror of Julia's DomainError thrown by `_assure_nonneg` (src/posterior/likelihood.jl:61-65)."""


# ----------------------------------------------------------------------------------------------
# A.1 kernels (KernelFunctions.jl: SqExponentialKernel, Matern32Kernel, Matern52Kernel with
# `with_lengthscale` -> ARDTransform(1 ./ lambda)); used by the reference at
# examples/simple/main.jl:36, examples/interactive/toy_problem.jl:58, src/acquisitions/immd.jl:44-45.
# ----------------------------------------------------------------------------------------------
def kernel_from_r2(kernel_id: int, r2: np.ndarray) -> np.ndarray:
    """k as a function of the squared scaled distance r2 = ||x/lam - x'/lam||^2."""
    if kernel_id == SE:
        return np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    if kernel_id == MATERN32:
        s = math.sqrt(3.0) * r
        return (1.0 + s) * np.exp(-s)
    if kernel_id == MATERN52:
        s = math.sqrt(5.0) * r
        return (1.0 + s + 5.0 * r2 / 3.0) * np.exp(-s)
    raise ValueError(f"unknown kernel id {kernel_id}")


def scaled_sqdist(A: np.ndarray, B: np.ndarray, lam: np.ndarray, gemm_trick: bool = False) -> np.ndarray:
    """r2[i, n] = sum_k ((A[k,i] - B[k,n]) / lam[k])^2.   A: d x M, B: d x N  ->  M x N.

    ``gemm_trick=True`` uses ||a||^2 + ||b||^2 - 2 a.b like Distances.pairwise(SqEuclidean) does in the
    reference's dependency chain; the default is the direct sum, which is
    the better-conditioned statement used for parity."""
    Au = (A / lam[:, None]).T  # M x d
    Bu = (B / lam[:, None]).T  # N x d
    if gemm_trick:
        r2 = (Au * Au).sum(1)[:, None] + (Bu * Bu).sum(1)[None, :] - 2.0 * (Au @ Bu.T)
        return np.maximum(r2, 0.0)
    r2 = np.zeros((Au.shape[0], Bu.shape[0]))
    for k in range(Au.shape[1]):
        diff = Au[:, k][:, None] - Bu[:, k][None, :]
        r2 += diff * diff
    return r2


def cross_kernel(kernel_id: int, A: np.ndarray, B: np.ndarray, lam: np.ndarray, gemm_trick: bool = False) -> np.ndarray:
    """Unscaled cross-kernel matrix k(A, B) (M x N); KernelFunctions.kernelmatrix(k, A, B)."""
    return kernel_from_r2(kernel_id, scaled_sqdist(A, B, lam, gemm_trick))


# ----------------------------------------------------------------------------------------------
# A.2 GP fit / predict: BOSS.model_posterior (called at src/posterior/log_posterior.jl:81,115,148)
# -> AbstractGPs.posterior semantics, independent GP per output (src/posterior/likelihood.jl:32-33,
# src/likelihoods/combined.jl:38,59,76).  Hyper-parameter layout lambda d x p, alpha p, sigma p
# (src/subset_problem.jl:38-44); data X d x N, Y p x N (src/subset_problem.jl:46-50).
# ----------------------------------------------------------------------------------------------
@dataclass
class GPOpts:
    jitter: float = 0.0       # extra diagonal variance (AbstractGPs default FiniteGP jitter is 1e-18; App. B item 3)
    pred_noise: bool = False  # include sigma_j^2 in the predictive variance (App. B item 2)
    clamp_var: bool = True    # clamp negative predictive variances to 0 before sqrt (App. B item 4)


@dataclass
class GPFit:
    kernel_id: int
    X: np.ndarray        # d x N
    lam: np.ndarray      # d x p
    alpha: np.ndarray    # p
    sigma: np.ndarray    # p
    L: np.ndarray        # p x N x N lower Cholesky factors of K_j
    a: np.ndarray        # N x p   a_j = K_j^{-1} (Y_j - m_j(X))
    opts: GPOpts = field(default_factory=GPOpts)

    @property
    def d(self): return self.X.shape[0]
    @property
    def N(self): return self.X.shape[1]
    @property
    def p(self): return self.alpha.shape[0]


def gp_fit(kernel_id, X, Y, lam, alpha, sigma, mean_at_X=None, opts: Optional[GPOpts] = None) -> GPFit:
    """K_j = alpha_j^2 k(X/lam_j) + (sigma_j^2 + jitter) I ;  L_j L_j^T = K_j ;  a_j = K_j^{-1}(Y_j - m_j(X))."""
    opts = opts or GPOpts()
    X = np.asarray(X, float); Y = np.asarray(Y, float)
    lam = np.asarray(lam, float).reshape(X.shape[0], -1)
    alpha = np.asarray(alpha, float); sigma = np.asarray(sigma, float)
    p, N = Y.shape
    L = np.empty((p, N, N)); a = np.empty((N, p))
    for j in range(p):
        K = alpha[j] ** 2 * cross_kernel(kernel_id, X, X, lam[:, j])
        K[np.diag_indices(N)] += sigma[j] ** 2 + opts.jitter
        Lj = cholesky(K, lower=True)
        resid = Y[j] - (0.0 if mean_at_X is None else np.asarray(mean_at_X)[j])
        a[:, j] = cho_solve((Lj, True), resid)
        L[j] = Lj
    return GPFit(kernel_id, X, lam, alpha, sigma, L, a, opts)


def gp_predict(fit: GPFit, Xq, mean_at_Xq=None, want_var=True, gemm_trick=False):
    """BOSS.mean / BOSS.var / mean_and_var (model_post, X) -> p x M  (call sites: src/posterior/likelihood.jl:8,12,21,25,
    src/likelihoods/normal.jl:37,43,60,68).  mu_j = m_j + k*.a_j ;  var_j = alpha_j^2 - ||L_j^{-1} k*||^2 (+sigma_j^2)."""
    Xq = np.asarray(Xq, float)
    M = Xq.shape[1]
    mu = np.empty((fit.p, M)); var = np.empty((fit.p, M)) if want_var else None
    for j in range(fit.p):
        Ks = fit.alpha[j] ** 2 * cross_kernel(fit.kernel_id, Xq, fit.X, fit.lam[:, j], gemm_trick)  # M x N
        mu[j] = Ks @ fit.a[:, j]
        if mean_at_Xq is not None:
            mu[j] += np.asarray(mean_at_Xq)[j]
        if want_var:
            V = solve_triangular(fit.L[j], Ks.T, lower=True, check_finite=False)  # N x M  (the TRSM of SURVEY 3.2)
            v = fit.alpha[j] ** 2 - np.einsum("nm,nm->m", V, V)
            if fit.opts.pred_noise:
                v = v + fit.sigma[j] ** 2
            if fit.opts.clamp_var:
                v = np.maximum(v, 0.0)
            var[j] = v
    return (mu, var) if want_var else mu


# ----------------------------------------------------------------------------------------------
# A.3 Normal log-pdf: Distributions.logpdf(Normal(mu, s), z) = StatsFuns.normlogpdf
#     = -(((z - mu)/s)^2 + log(2 pi))/2 - log(s)
# ----------------------------------------------------------------------------------------------
def normlogpdf(mu, s, z):
    zval = (z - mu) / s
    return -(zval * zval + LOG2PI) / 2.0 - np.log(s)


# ----------------------------------------------------------------------------------------------
# A.4 NormalLikelihood: src/likelihoods/normal.jl
# ----------------------------------------------------------------------------------------------
@dataclass
class NormalLikelihood:
    """src/likelihoods/normal.jl:16-19"""
    z_obs: np.ndarray
    std_obs: np.ndarray

    def __post_init__(self):
        self.z_obs = np.asarray(self.z_obs, float); self.std_obs = np.asarray(self.std_obs, float)


def loglike_marginal(like: NormalLikelihood, Y):
    """src/likelihoods/normal.jl:28-30 (vector) and the column-broadcast fallback src/types/likelihood.jl:53-54.
    Y: p (vector) or p x M -> same shape."""
    Y = np.asarray(Y, float)
    if Y.ndim == 1:
        return normlogpdf(Y, like.std_obs, like.z_obs)
    return normlogpdf(Y, like.std_obs[:, None], like.z_obs[:, None])


def loglike(like: NormalLikelihood, Y):
    """src/types/likelihood.jl:66-67,70-71: sum of the marginal terms (scalar, or length-M vector)."""
    return loglike_marginal(like, Y).sum(axis=0)


def log_marginal_likelihood_mean(like: NormalLikelihood, mu, std):
    """src/likelihoods/normal.jl:32-53: logpdf(Normal(mu, sqrt(std_obs^2 + std_y^2)), z_obs); p or p x M."""
    mu = np.asarray(mu, float); std = np.asarray(std, float)
    so = like.std_obs if mu.ndim == 1 else like.std_obs[:, None]
    z = like.z_obs if mu.ndim == 1 else like.z_obs[:, None]
    s = np.sqrt(so ** 2 + std ** 2)
    return normlogpdf(mu, s, z)


def log_likelihood_mean(like, mu, std):
    """src/posterior/likelihood.jl:34-43: sum over output dims of the marginal expectation."""
    return log_marginal_likelihood_mean(like, mu, std).sum(axis=0)


def log_sq_likelihood_mean(like: NormalLikelihood, mu, std):
    """src/likelihoods/normal.jl:55-79: log E[l^2] = -sum log(2 sqrt(pi) so) + sum logN(z; mu, sqrt((so^2 + 2 std^2)/2))."""
    mu = np.asarray(mu, float); std = np.asarray(std, float)
    so = like.std_obs if mu.ndim == 1 else like.std_obs[:, None]
    z = like.z_obs if mu.ndim == 1 else like.z_obs[:, None]
    s = np.sqrt((so ** 2 + 2.0 * std ** 2) / 2.0)
    log_C = -np.sum(np.log(2.0 * math.sqrt(math.pi) * like.std_obs))
    return log_C + normlogpdf(mu, s, z).sum(axis=0)


def assure_nonneg(v):
    """src/posterior/likelihood.jl:61-65, vectorised: v>=0 keep; -1e-8<=v<0 -> 0; v<-1e-8 -> DomainError."""
    v = np.asarray(v, float)
    bad = v < -MAX_NEG_VAR
    if np.any(bad):
        i = int(np.argmax(bad.reshape(-1)))
        raise DomainError(f"Expected a non-negative value, but got {v.reshape(-1)[i]} (index {i}).")
    return np.where(v >= 0, v, 0.0)


def log_likelihood_variance(like, mu, std):
    """src/posterior/likelihood.jl:47-59: log( exp(log E[l^2]) - exp(2 log E[l]) ), difference in linear space."""
    log_sqL = log_sq_likelihood_mean(like, mu, std)
    log_L = log_likelihood_mean(like, mu, std)
    like_var = assure_nonneg(np.exp(log_sqL) - np.exp(2.0 * log_L))
    with np.errstate(divide="ignore"):
        return np.log(like_var)


# ----------------------------------------------------------------------------------------------
# The other closed-form likelihoods of the Gaussian family.  Each exposes the same three reductions as NormalLikelihood:
#   loglike(mu)                         plug-in log-likelihood      (scalar / length M)
#   log_likelihood_mean(mu, std)        log E[l]
#   log_likelihood_variance(mu, std)    log V[l]   (via log E[l^2] and the clamp of src/posterior/likelihood.jl:47-65)
# through `like_terms(like, mu, var)` below.
# ----------------------------------------------------------------------------------------------
@dataclass
class LogNormalLikelihood:
    """src/likelihoods/lognormal.jl:26-91"""
    log_z_obs: np.ndarray
    CV: np.ndarray

    def __post_init__(self):
        self.log_z_obs = np.asarray(self.log_z_obs, float); self.CV = np.asarray(self.CV, float)
        self.z_obs = np.exp(self.log_z_obs)                       # :34
        self.sigma_log = np.sqrt(np.log(1.0 + self.CV ** 2))      # :45


@dataclass
class NormalDiffLikelihood:
    """src/likelihoods/normal_diff.jl:17-73"""
    std_obs: np.ndarray

    def __post_init__(self):
        self.std_obs = np.asarray(self.std_obs, float)


@dataclass
class NormalSumLikelihood:
    """src/likelihoods/normal_sum.jl:18-115"""
    sum_lengths: Sequence[int]
    z_obs: np.ndarray
    std_obs: np.ndarray

    def __post_init__(self):
        self.z_obs = np.asarray(self.z_obs, float); self.std_obs = np.asarray(self.std_obs, float)


@dataclass
class ExpLikelihood:
    """src/likelihoods/exp.jl:7-63"""


def _lognormal_logpdf(mu_log, s, z):
    """Distributions.logpdf(LogNormal(mu, s), z) = normlogpdf(mu, s, log z) - log z"""
    return normlogpdf(mu_log, s, np.log(z)) - np.log(z)


def _indexed_sum(Y, sum_lengths):
    """_indexed_sum (src/likelihoods/normal_sum.jl:29-40) over the leading axis."""
    out, idx = [], 0
    for n in sum_lengths:
        out.append(Y[idx:idx + n].sum(axis=0)); idx += n
    return np.stack(out)


def like_terms(like, mu, var):
    """mu, var: p (vector) or p x M.  Returns dict(loglike_marginal, loglike, log_marginal_mean, log_like_mean, log_sq_like_mean,
    log_like_var) restating the per-likelihood reference code cited in the class docstrings."""
    mu = np.asarray(mu, float); var = np.zeros_like(mu) if var is None else np.asarray(var, float)
    col = (lambda v: v) if mu.ndim == 1 else (lambda v: np.asarray(v)[:, None])
    if isinstance(like, ExpLikelihood):                                     # exp.jl:9-16, 19-33, 50-63
        m, v = mu[0], var[0]
        with np.errstate(divide="ignore"):
            return {"loglike_marginal": mu, "loglike": m, "log_marginal_mean": mu + 0.5 * var, "log_like_mean": m + 0.5 * v,
                    "log_sq_like_mean": 2 * m + 2 * v, "log_like_var": 2 * (m + v) + np.log(1 - np.exp(-v))}
    if isinstance(like, LogNormalLikelihood):
        z, s0 = col(like.z_obs), col(like.sigma_log)
        mu_log = mu - s0 ** 2 / 2                                           # lognormal.jl:44
        lmarg = _lognormal_logpdf(mu_log, s0, z)                            # :47-50
        mmarg = _lognormal_logpdf(mu_log, np.sqrt(s0 ** 2 + var), z)        # :56-61
        log_C = -np.sum(np.log(2 * math.sqrt(math.pi) * like.sigma_log * like.z_obs))              # :83
        lsq = log_C + _lognormal_logpdf(mu_log, np.sqrt((s0 ** 2 + 2 * var) / 2), z).sum(axis=0)   # :81-84
    else:
        if isinstance(like, NormalSumLikelihood):
            mu = _indexed_sum(mu, like.sum_lengths); var = _indexed_sum(var, like.sum_lengths)     # normal_sum.jl:42-46
            z, s0 = col(like.z_obs), col(like.std_obs)
        elif isinstance(like, NormalDiffLikelihood):
            z, s0 = col(np.zeros_like(like.std_obs)), col(like.std_obs)                            # normal_diff.jl:21-23
        else:
            z, s0 = col(like.z_obs), col(like.std_obs)
        lmarg = normlogpdf(mu, s0, z)
        mmarg = normlogpdf(mu, np.sqrt(s0 ** 2 + var), z)
        log_C = -np.sum(np.log(2 * math.sqrt(math.pi) * np.ravel(s0)))
        lsq = log_C + normlogpdf(mu, np.sqrt((s0 ** 2 + 2 * var) / 2), z).sum(axis=0)
    lmean = mmarg.sum(axis=0)
    v = assure_nonneg(np.exp(lsq) - np.exp(2.0 * lmean))                    # src/posterior/likelihood.jl:55-57
    with np.errstate(divide="ignore"):
        lvar = np.log(v)
    return {"loglike_marginal": lmarg, "loglike": lmarg.sum(axis=0), "log_marginal_mean": mmarg, "log_like_mean": lmean,
            "log_sq_like_mean": lsq, "log_like_var": lvar}


# ----------------------------------------------------------------------------------------------
# A.5 priors and posterior closures: src/posterior/log_posterior.jl
# ----------------------------------------------------------------------------------------------
UNIFORM, NORMAL, TRUNCNORMAL = 0, 1, 2


@dataclass
class ProductPrior:
    """Product of independent univariate priors, one per parameter (examples/simple/toy_problem.jl:31 uses
    truncated Normals, examples/interactive/toy_problem.jl:43 Uniforms, docs/src/example.md:52-55 Normals).
    kind[k] in {UNIFORM, NORMAL, TRUNCNORMAL}; params[k] = (a, b) | (m, s) | (m, s, a, b)."""
    kind: Sequence[int]
    params: Sequence[Sequence[float]]

    @property
    def d(self): return len(self.kind)


def truncnormal_logtp(m, s, a, b):
    """log(Phi((b-m)/s) - Phi((a-m)/s)) -- the normaliser Distributions.Truncated stores as `logtp`."""
    return math.log(float(ndtr((b - m) / s)) - float(ndtr((a - m) / s)))


def log_prior(prior: ProductPrior, X):
    """_log_prior: logpdf.(Ref(prior), eachcol(X)) (src/posterior/log_posterior.jl:228-229); X d x M -> M; -Inf outside support."""
    X = np.asarray(X, float)
    vec = X.ndim == 1
    if vec:
        X = X[:, None]
    lp = np.zeros(X.shape[1])
    for k in range(prior.d):
        x = X[k]
        kind, prm = prior.kind[k], prior.params[k]
        if kind == UNIFORM:
            a, b = prm[:2]
            t = np.where((x >= a) & (x <= b), -math.log(b - a), -np.inf)
        elif kind == NORMAL:
            m, s = prm[:2]
            t = normlogpdf(m, s, x)
        elif kind == TRUNCNORMAL:
            m, s, a, b = prm[:4]
            t = np.where((x >= a) & (x <= b), normlogpdf(m, s, x) - truncnormal_logtp(m, s, a, b), -np.inf)
        else:
            raise ValueError(kind)
        lp = lp + t
    return lp[0] if vec else lp


def posterior_eval(fit: GPFit, like: NormalLikelihood, prior: ProductPrior, Xq, mean_at_Xq=None, gemm_trick=False):
    """All three log-closures of src/posterior/log_posterior.jl on one batch:
       log_approx_posterior :15-21  = lp + sum_j logN(z_j; mu_j, so_j)              (plug-in, src/posterior/likelihood.jl:19-29)
       log_posterior_mean   :35-41  = lp + log E[l]
       log_posterior_variance :55-61 = 2 lp + log V[l]
    Returns dict with those three length-M vectors plus mu, var (p x M)."""
    mu, var = gp_predict(fit, Xq, mean_at_Xq, True, gemm_trick)
    std = np.sqrt(var)
    lp = log_prior(prior, Xq)
    out = {
        "mu": mu, "var": var, "log_prior": lp,
        "log_approx_post": lp + loglike(like, mu),
        "log_post_mean": lp + log_likelihood_mean(like, mu, std),
        "log_post_var": 2.0 * lp + log_likelihood_variance(like, mu, std),
    }
    if isinstance(like, NormalLikelihood):   # the two terms of V = exp(log E[l^2]) - exp(2 log E[l]): the tests bound the cancellation with them
        out["log_like_mean"] = log_likelihood_mean(like, mu, std)
        out["log_sq_like_mean"] = log_sq_likelihood_mean(like, mu, std)
    return out


def posterior_eval_bi(fits: Sequence[GPFit], like: NormalLikelihood, prior: ProductPrior, Xq):
    """Bayesian-inference (MultiFittedParams) closures, src/posterior/log_posterior.jl:86-95,120-129,153-162:
    log.(mapreduce(f -> exp.(f(x)), .+, fs) ./ sample_count) over the per-sample likelihood closures (linear-space average,
    as the reference does), then the prior added as in :19,:39,:59."""
    lp = log_prior(prior, Xq)
    acc = {"a": 0.0, "m": 0.0, "v": 0.0}
    for fit in fits:
        mu, var = gp_predict(fit, Xq)
        std = np.sqrt(var)
        acc["a"] = acc["a"] + np.exp(loglike(like, mu))
        acc["m"] = acc["m"] + np.exp(log_likelihood_mean(like, mu, std))
        acc["v"] = acc["v"] + np.exp(log_likelihood_variance(like, mu, std))
    S = len(fits)
    with np.errstate(divide="ignore"):
        return {"log_prior": lp, "log_approx_post": lp + np.log(acc["a"] / S), "log_post_mean": lp + np.log(acc["m"] / S),
                "log_post_var": 2.0 * lp + np.log(acc["v"] / S)}


def maxvar_argmax(log_post_var: np.ndarray, log_space: bool):
    """MaxVar (src/acquisitions/maxvar.jl:9-11 -> posterior.jl:141: exp.(log V)) or LogMaxVar
    (src/acquisitions/logmaxvar.jl:13-15) evaluated on a candidate set, then Julia `argmax` semantics of BOSS's
    SamplingAM/GridAM (first maximal index).  Returns (0-based index, value)."""
    vals = log_post_var if log_space else np.exp(log_post_var)
    i = int(np.argmax(vals))  # numpy argmax also returns the first maximal index
    return i, float(vals[i])


# ----------------------------------------------------------------------------------------------
# marginal-integration sweep: _compute_marginals_int, ext/CairoExt.jl:146-246 (un-normalised means)
# ----------------------------------------------------------------------------------------------
def marginals_int(f, lhc: np.ndarray, lb, ub, grid_size: int, diag=True, pairs=True):
    """f: matrix closure X (d x G) -> length-G values (e.g. exp(log_post_mean)).  Restates the reference's loops: overwrite
    row(s) of the LHC grid with the range value(s), ys[i] = mean(f(grid)), restore.  Returns (diag gs x d, pairs gs x gs x npairs)."""
    grid = np.array(lhc, dtype=float)
    d = grid.shape[0]
    axes = [np.linspace(lb[k], ub[k], grid_size) for k in range(d)]
    out_d = np.zeros((grid_size, d)) if diag else None
    out_p = np.zeros((grid_size, grid_size, d * (d - 1) // 2)) if pairs else None
    if diag:
        for k in range(d):
            tmp = grid[k].copy()
            for i, x_ in enumerate(axes[k]):
                grid[k] = x_
                out_d[i, k] = np.mean(f(grid))
            grid[k] = tmp
    if pairs:
        pi = 0
        for a in range(d):
            for b in range(a + 1, d):
                ta, tb = grid[a].copy(), grid[b].copy()
                for ib, xb in enumerate(axes[b]):
                    for ia, xa in enumerate(axes[a]):
                        grid[a] = xa; grid[b] = xb
                        out_p[ia, ib, pi] = np.mean(f(grid))
                grid[a] = ta; grid[b] = tb
                pi += 1
    return out_d, out_p


# ----------------------------------------------------------------------------------------------
# A.6 evidence: src/posterior/evidence.jl:19-24
# ----------------------------------------------------------------------------------------------
def evidence(log_post_vals: np.ndarray, log_prior_vals: np.ndarray) -> float:
    """mean_i post(x_i)/pdf(prior, x_i) given the log-posterior and log-prior at prior samples x_i."""
    return float(np.mean(np.exp(log_post_vals) / np.exp(log_prior_vals)))


# ----------------------------------------------------------------------------------------------
# A.7 MMD: src/metrics/mmd.jl:23-28  (biased V-statistic, diagonals included)
# ----------------------------------------------------------------------------------------------
def calc_mmd(kernel_id: int, lam: np.ndarray, A: np.ndarray, B: np.ndarray, block: int = 2048, gemm_trick=False):
    """mean(k(A,A)) + mean(k(B,B)) - 2 mean(k(A,B));  A: d x n, B: d x m.  Tiled so that it runs for large n.
    Returns (mmd2, sum_AA, sum_BB, sum_AB)."""
    A = np.asarray(A, float); B = np.asarray(B, float); lam = np.asarray(lam, float)

    def ksum(P, Q):
        tot = 0.0
        for i0 in range(0, P.shape[1], block):
            for j0 in range(0, Q.shape[1], block):
                tot += float(cross_kernel(kernel_id, P[:, i0:i0 + block], Q[:, j0:j0 + block], lam, gemm_trick).sum())
        return tot

    n, m = A.shape[1], B.shape[1]
    sAA, sBB, sAB = ksum(A, A), ksum(B, B), ksum(A, B)
    return sAA / (n * n) + sBB / (m * m) - 2.0 * sAB / (n * m), sAA, sBB, sAB


# ----------------------------------------------------------------------------------------------
# A.8 TV: src/metrics/tv.jl:59-73, src/utils/utils.jl:7-15
# ----------------------------------------------------------------------------------------------
def logsumexp(x):
    """src/utils/utils.jl:7-11"""
    x = np.asarray(x, float)
    x_max = np.max(x)
    if np.isinf(x_max):
        return float(x_max)
    return float(x_max + np.log(np.sum(np.exp(x - x_max))))


def logmeanexp(x):
    """src/utils/utils.jl:13-15"""
    return logsumexp(x) - math.log(len(x))


def calc_tv(log_ws, approx_logvals, true_logvals):
    """src/metrics/tv.jl:59-73"""
    log_ws = np.asarray(log_ws, float); a = np.asarray(approx_logvals, float); t = np.asarray(true_logvals, float)
    ev_a = logmeanexp(log_ws + a)
    ev_t = logmeanexp(log_ws + t)
    a_n = np.full_like(a, -np.inf) if np.isinf(ev_a) else a - ev_a
    t_n = np.full_like(t, -np.inf) if np.isinf(ev_t) else t - ev_t
    diffs = np.abs(np.exp(a_n) - np.exp(t_n))
    with np.errstate(divide="ignore"):
        return 0.5 * math.exp(logmeanexp(log_ws + np.log(diffs)))


# ----------------------------------------------------------------------------------------------
# CPU baseline ("the reference's CPU path restated"): same algorithmic
# structure as the reference -- pairwise-distance K*, GEMV mean, BLAS-3 TRSM against L, column sum
# of squares, vectorised epilogue; the GP predictive is evaluated TWICE per variance call as
# src/posterior/likelihood.jl:48-49 -> src/likelihoods/normal.jl:43,68 does (dedup=False).
# ----------------------------------------------------------------------------------------------
def cpu_baseline_posterior_maxvar(fit: GPFit, like, prior, Xq, chunk=65536, dedup=False):
    """approx_posterior + posterior_mean + MaxVar over Xq in chunks; returns (log_approx, log_mean, log_var, argmax)."""
    M = Xq.shape[1]
    out_a = np.empty(M); out_m = np.empty(M); out_v = np.empty(M)
    for c0 in range(0, M, chunk):
        Xc = Xq[:, c0:c0 + chunk]
        lp = log_prior(prior, Xc)
        mu = gp_predict(fit, Xc, want_var=False, gemm_trick=True)               # approx_posterior: BOSS.mean
        out_a[c0:c0 + chunk] = lp + loglike(like, mu)
        mu1, var1 = gp_predict(fit, Xc, gemm_trick=True)                        # predict #1 (normal.jl:43)
        std1 = np.sqrt(var1)
        log_L = log_likelihood_mean(like, mu1, std1)
        if dedup:
            mu2, std2 = mu1, std1
        else:
            mu2, var2 = gp_predict(fit, Xc, gemm_trick=True)                    # predict #2 (normal.jl:68)
            std2 = np.sqrt(var2)
        log_sq = log_sq_likelihood_mean(like, mu2, std2)
        v = assure_nonneg(np.exp(log_sq) - np.exp(2.0 * log_L))
        with np.errstate(divide="ignore"):
            out_v[c0:c0 + chunk] = 2.0 * lp + np.log(v)
        out_m[c0:c0 + chunk] = lp + log_L
    return out_a, out_m, out_v, int(np.argmax(out_v))


# ------------------------------------------------------------------------------------------------ AMIS inner loop, confidence sets
# (test infrastructure: restates /root/reference/src/samplers/amis/amis_method.jl:26-134, proposal_distributions/normal.jl:75-100,
#  src/utils/confidence_set.jl:19-81, src/sampling.jl:55-57 and StatsBase's weighted quantile, which Distributions.weights feeds)
def mvnormal_logpdf(mu, Sigma, X):
    """logpdf(MvNormal(mu, Sigma), x) per column of X (d x M)."""
    mu = np.asarray(mu, float); Sigma = np.asarray(Sigma, float); X = np.asarray(X, float).reshape(len(mu), -1)
    Lc = np.linalg.cholesky(Sigma)
    z = solve_triangular(Lc, X - mu[:, None], lower=True)
    return -0.5 * np.sum(z * z, axis=0) - np.sum(np.log(np.diag(Lc))) - 0.5 * len(mu) * math.log(2.0 * math.pi)


def exp_weights(log_ws):
    """amis_method.jl:126-134"""
    log_ws = np.asarray(log_ws, float)
    M = np.max(log_ws)
    if M == -np.inf:
        raise RuntimeError("Sampling failed due to numerical issues with sample weights.")
    ws = np.exp(log_ws - M)
    return ws / np.sum(ws)


def get_weights(log_ws):
    """amis_method.jl:109-124"""
    log_ws = np.asarray(log_ws, float)
    if np.any(np.isnan(log_ws)) or np.any(log_ws == np.inf) or np.all(log_ws == -np.inf):
        raise RuntimeError("AMIS: Numerical issues with sample weights.")
    return exp_weights(log_ws)


def estimate_parameters(xs, ws):
    """estimate_parameters!(::NormalProposal, xs, ws) -- normal.jl:75-100; returns (mu, Sigma) or None when the estimate is unusable."""
    xs = np.asarray(xs, float); ws = np.asarray(ws, float)
    V1, V2 = np.sum(ws), np.sum(ws ** 2)
    mu = (xs * ws[None, :]).sum(axis=1) / V1
    dl = xs - mu[:, None]
    Sigma = (dl * ws[None, :]) @ dl.T / (V1 - V2 / V1)
    Sigma = 0.5 * (Sigma + Sigma.T)
    if np.any(np.isinf(Sigma)) or np.any(np.isnan(Sigma)):
        return None
    try:
        np.linalg.cholesky(Sigma)
    except np.linalg.LinAlgError:
        return None
    return mu, Sigma


def amis_replay(xs, log_P, N, q0, init_q=None):
    """Replays (amis::AMIS)(log_pi, q, fitter) (amis_method.jl:26-83) on GIVEN draws xs (d x N(T+1), iteration-major) and their
    log_pi values: returns the proposals fitted at every iteration, Delta, log_Omega and the final weights of iterations 2..T+1."""
    xs = np.asarray(xs, float); log_P = np.asarray(log_P, float)
    T1 = xs.shape[1] // N
    qs = [tuple(np.array(a, float) for a in q0) for _ in range(T1)]
    if init_q is not None:
        qs[0] = tuple(np.array(a, float) for a in init_q)
    Delta = np.zeros(N * T1); log_O = np.zeros(N * T1)
    sl = lambda t0, t1: slice(t0 * N, t1 * N)  # noqa: E731
    Delta[sl(0, 1)] = np.exp(mvnormal_logpdf(*qs[0], xs[:, sl(0, 1)]))
    log_O[sl(0, 1)] = log_P[sl(0, 1)] - np.log(Delta[sl(0, 1)])
    for t in range(1, T1):
        est = estimate_parameters(xs[:, sl(0, t)], get_weights(log_O[sl(0, t)]))
        if est is not None:
            qs[t] = est
        for i in range(t + 1):
            Delta[sl(t, t + 1)] += np.exp(mvnormal_logpdf(*qs[i], xs[:, sl(t, t + 1)]))
        log_O[sl(t, t + 1)] = log_P[sl(t, t + 1)] - np.log(Delta[sl(t, t + 1)] / (t + 1))
        for i in range(t):
            Delta[sl(i, i + 1)] += np.exp(mvnormal_logpdf(*qs[t], xs[:, sl(i, i + 1)]))
            log_O[sl(i, i + 1)] = log_P[sl(i, i + 1)] - np.log(Delta[sl(i, i + 1)] / (t + 1))
    return qs, Delta, log_O, (get_weights(log_O[sl(1, T1)]) if T1 > 1 else None)


def weighted_quantile(v, w, p):
    """StatsBase.quantile(v, w::Weights, p) restated loop for loop (zero weights removed, sorted by value)."""
    v = np.asarray(v, float); w = np.asarray(w, float)
    wsum = float(np.sum(w))
    vw = sorted((float(a), float(b)) for a, b in zip(v, w) if b != 0)
    if any(math.isnan(a) for a in v):
        return float("nan")
    out = vw[-1][0]
    Sk = Skold = 0.0; vk = vkold = 0.0; k = 0
    w1 = vw[0][1]
    h = p * (wsum - w1) + w1
    while Sk <= h:
        k += 1
        if k > len(vw):
            return out
        Skold, vkold = Sk, vk
        vk, wk = vw[k - 1]
        Sk += wk
    return vkold + (h - Skold) / (Sk - Skold) * (vk - vkold)


def find_cutoff(vals, ws, q):
    """confidence_set.jl:23-27 on precomputed target_pdf values"""
    return weighted_quantile(vals, ws, 1.0 - q)


def approx_cutoff_area(vals, ws, c):
    """confidence_set.jl:51-56"""
    ws = np.asarray(ws, float)
    wn = np.exp(np.log(ws) - np.log(np.sum(ws)))
    return float(np.sum(wn[np.asarray(vals) >= c]))


def set_iou(in_A, in_B, prior_pdf_vals):
    """confidence_set.jl:73-81"""
    ws = 1.0 / np.asarray(prior_pdf_vals, float)
    ws = ws / np.sum(ws)
    a, b = np.asarray(in_A, bool), np.asarray(in_B, bool)
    return float(np.sum(ws[a & b]) / np.sum(ws[a | b]))
