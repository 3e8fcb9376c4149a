"""50-digit mpmath restatement of the same formulas (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Independent of bosip_oracle.py on purpose: scalar loops, no NumPy linear algebra, so that the float64
oracle can be checked against it and so that tests/golden/*.json hold values good to the last float64
digit.  Only usable for tiny problems (N <= ~12, M <= ~16)."""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50

SE, MATERN32, MATERN52 = 0, 1, 2


def _kernel(kernel_id, r2):
    # SURVEY.md Appendix A.1 (KernelFunctions.jl semantics)
    if kernel_id == SE:
        return mp.exp(-r2 / 2)
    r = mp.sqrt(r2)
    if kernel_id == MATERN32:
        s = mp.sqrt(3) * r
        return (1 + s) * mp.exp(-s)
    s = mp.sqrt(5) * r
    return (1 + s + 5 * r2 / 3) * mp.exp(-s)


def _r2(x, y, lam):
    return sum(((mp.mpf(xi) - mp.mpf(yi)) / mp.mpf(li)) ** 2 for xi, yi, li in zip(x, y, lam))


def normlogpdf(mu, s, z):
    # SURVEY.md Appendix A.3
    zv = (mp.mpf(z) - mp.mpf(mu)) / mp.mpf(s)
    return -(zv * zv + mp.log(2 * mp.pi)) / 2 - mp.log(s)


def gp_predict(kernel_id, X, Y, lam, alpha, sigma, Xq, jitter=0.0, pred_noise=False):
    """X: list of N points (each length d); Y: p lists of N; lam: p lists of d; Xq: list of M points.
    Returns (mu[p][M], var[p][M]) -- SURVEY.md Appendix A.2."""
    N, p = len(X), len(Y)
    mu = [[None] * len(Xq) for _ in range(p)]
    var = [[None] * len(Xq) for _ in range(p)]
    for j in range(p):
        a2 = mp.mpf(alpha[j]) ** 2
        K = mp.matrix(N, N)
        for i in range(N):
            for k in range(N):
                K[i, k] = a2 * _kernel(kernel_id, _r2(X[i], X[k], lam[j]))
            K[i, i] += mp.mpf(sigma[j]) ** 2 + mp.mpf(jitter)
        y = mp.matrix([mp.mpf(v) for v in Y[j]])
        a = mp.lu_solve(K, y)
        for q, xq in enumerate(Xq):
            ks = mp.matrix([a2 * _kernel(kernel_id, _r2(xq, X[i], lam[j])) for i in range(N)])
            mu[j][q] = sum(ks[i] * a[i] for i in range(N))
            w = mp.lu_solve(K, ks)
            v = a2 - sum(ks[i] * w[i] for i in range(N))
            if pred_noise:
                v += mp.mpf(sigma[j]) ** 2
            var[j][q] = v
    return mu, var


def likelihood_terms(z_obs, std_obs, mu, var):
    """mu, var: length-p lists at ONE point.  Returns dict of the scalar quantities of SURVEY.md Appendix A.4."""
    p = len(z_obs)
    plug = [normlogpdf(mu[j], std_obs[j], z_obs[j]) for j in range(p)]
    lm = [normlogpdf(mu[j], mp.sqrt(mp.mpf(std_obs[j]) ** 2 + mp.mpf(var[j])), z_obs[j]) for j in range(p)]
    log_C = -sum(mp.log(2 * mp.sqrt(mp.pi) * mp.mpf(std_obs[j])) for j in range(p))
    lsq = log_C + sum(normlogpdf(mu[j], mp.sqrt((mp.mpf(std_obs[j]) ** 2 + 2 * mp.mpf(var[j])) / 2), z_obs[j]) for j in range(p))
    log_L = sum(lm)
    V = mp.exp(lsq) - mp.exp(2 * log_L)
    return {"loglike_marginal": plug, "loglike": sum(plug), "log_ml_mean": lm, "log_like_mean": log_L,
            "log_sq_like_mean": lsq, "like_var": V, "log_like_var": (mp.log(V) if V > 0 else mp.mpf("-inf"))}


def log_prior(kind, params, x):
    """Product prior log-pdf. kind: 0 uniform (a,b), 1 normal (m,s), 2 truncnormal (m,s,a,b)."""
    lp = mp.mpf(0)
    for k, xk in enumerate(x):
        xk = mp.mpf(xk)
        prm = [mp.mpf(v) for v in params[k]]
        if kind[k] == 0:
            a, b = prm[:2]
            if not (a <= xk <= b):
                return mp.mpf("-inf")
            lp -= mp.log(b - a)
        elif kind[k] == 1:
            m, s = prm[:2]
            lp += normlogpdf(m, s, xk)
        else:
            m, s, a, b = prm[:4]
            if not (a <= xk <= b):
                return mp.mpf("-inf")
            lp += normlogpdf(m, s, xk) - mp.log(mp.ncdf((b - m) / s) - mp.ncdf((a - m) / s))
    return lp


def calc_mmd(kernel_id, lam, A, B):
    """SURVEY.md Appendix A.7 (src/metrics/mmd.jl:23-28); A, B lists of points."""
    def mean_k(P, Q):
        return sum(_kernel(kernel_id, _r2(x, y, lam)) for x in P for y in Q) / (len(P) * len(Q))
    return mean_k(A, A) + mean_k(B, B) - 2 * mean_k(A, B)


def calc_tv(log_ws, a, t):
    """SURVEY.md Appendix A.8 (src/metrics/tv.jl:59-73) for finite inputs."""
    G = len(log_ws)
    ea = mp.log(sum(mp.exp(mp.mpf(w) + mp.mpf(v)) for w, v in zip(log_ws, a)) / G)
    et = mp.log(sum(mp.exp(mp.mpf(w) + mp.mpf(v)) for w, v in zip(log_ws, t)) / G)
    s = sum(mp.exp(mp.mpf(w)) * abs(mp.exp(mp.mpf(x) - ea) - mp.exp(mp.mpf(y) - et)) for w, x, y in zip(log_ws, a, t)) / G
    return s / 2
