"""Generate tests/golden/*.json with the 50-digit mpmath restatement (oracle/mp_oracle.py).

    python tests/golden/make_golden.py

The reference (Julia, BOSS.jl un-vendored) cannot be executed in this image, so these are NOT outputs of the reference
itself: (1) `likelihood_fixture.json` re-evaluates the reference's own test fixture
(/root/reference/test/unit/test/posterior/likelihood.jl:1-22; the identities it asserts at :24-136 are re-asserted in
tests/test_oracle_golden.py), (2) the rest are high-precision evaluations of the published formulas
on tiny seeded problems.  Values are stored as float64 (repr round-trips) plus 25-digit strings."""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import mp_oracle as mpo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def f(x):
    return float(x) if x != mp.mpf("-inf") else float("-inf")


def s(x):
    return mp.nstr(x, 25)


def likelihood_fixture():
    z, so, mu = [1.0, 2.0, 3.0], [0.5, 1.0, 2.0], [1.1, 2.2, 3.3]
    out = {"z_obs": z, "std_obs": so, "mean": mu, "cases": {}}
    for name, var in (("var_zero", [0.0, 0.0, 0.0]), ("var_nonzero", [0.25, 1.0, 4.0])):
        t = mpo.likelihood_terms(z, so, mu, var)
        out["cases"][name] = {
            "var": var,
            "loglike_marginal": [f(v) for v in t["loglike_marginal"]], "loglike": f(t["loglike"]),
            "log_ml_mean": [f(v) for v in t["log_ml_mean"]], "log_like_mean": f(t["log_like_mean"]),
            "log_sq_like_mean": f(t["log_sq_like_mean"]), "like_var": f(t["like_var"]), "log_like_var": f(t["log_like_var"]),
            "str": {"log_like_mean": s(t["log_like_mean"]), "log_sq_like_mean": s(t["log_sq_like_mean"]), "like_var": s(t["like_var"])},
        }
    return out


def gp_small():
    rng = np.random.Generator(np.random.PCG64(20261019))
    cases = []
    for kernel_id in (0, 1, 2):
        d, p, N, M = 2, 2, 6, 5
        X = rng.uniform(-5, 5, size=(d, N))
        Y = np.sin(X).sum(axis=0)[None, :] * np.array([[1.0], [0.5]]) + 0.1 * rng.standard_normal((p, N))
        lam = rng.uniform(1.0, 4.0, size=(d, p))
        alpha = np.array([1.3, 0.7]); sigma = np.array([0.1, 0.05])
        Xq = rng.uniform(-5, 5, size=(d, M))
        Xq[:, 0] = X[:, 2] + 1e-3           # close to a training point: small variance, cancellation
        Xq[:, 1] = [5.5, 0.0]               # outside the truncated/uniform support
        z_obs = [0.3, -0.2]; std_obs = [0.2, 0.4]
        Xl = [list(X[:, n]) for n in range(N)]; Yl = [list(Y[j]) for j in range(p)]
        laml = [list(lam[:, j]) for j in range(p)]; Xql = [list(Xq[:, i]) for i in range(M)]
        mu, var = mpo.gp_predict(kernel_id, Xl, Yl, laml, list(alpha), list(sigma), Xql)
        priors = {
            "truncnormal": ([2, 2], [[0.0, 5.0 / 3.0, -5.0, 5.0]] * 2),
            "uniform": ([0, 0], [[-5.0, 5.0, 0, 0], [-6.0, 6.0, 0, 0]]),
            "normal": ([1, 1], [[0.5, 2.0, 0, 0], [-0.5, 3.0, 0, 0]]),
        }
        pts = []
        for i in range(M):
            t = mpo.likelihood_terms(z_obs, std_obs, [mu[j][i] for j in range(p)], [var[j][i] for j in range(p)])
            lps = {k: mpo.log_prior(kd, prm, Xql[i]) for k, (kd, prm) in priors.items()}
            pts.append({"mu": [f(mu[j][i]) for j in range(p)], "var": [f(var[j][i]) for j in range(p)],
                        "loglike": f(t["loglike"]), "log_like_mean": f(t["log_like_mean"]),
                        "log_sq_like_mean": f(t["log_sq_like_mean"]), "like_var": f(t["like_var"]),
                        "log_like_var": f(t["log_like_var"]), "log_prior": {k: f(v) for k, v in lps.items()}})
        cases.append({"kernel_id": kernel_id, "X": X.tolist(), "Y": Y.tolist(), "lam": lam.tolist(), "alpha": alpha.tolist(),
                      "sigma": sigma.tolist(), "Xq": Xq.tolist(), "z_obs": z_obs, "std_obs": std_obs,
                      "priors": {k: {"kind": kd, "params": prm} for k, (kd, prm) in priors.items()}, "points": pts})
    return {"cases": cases}


def metrics_small():
    rng = np.random.Generator(np.random.PCG64(20261020))
    out = {"mmd": [], "tv": []}
    for kernel_id in (0, 1, 2):
        d, n, m = 3, 9, 7
        A = rng.standard_normal((d, n)); B = 0.3 + 1.2 * rng.standard_normal((d, m))
        lam = [0.8, 1.5, 2.2]
        v = mpo.calc_mmd(kernel_id, lam, [list(A[:, i]) for i in range(n)], [list(B[:, i]) for i in range(m)])
        out["mmd"].append({"kernel_id": kernel_id, "lam": lam, "A": A.tolist(), "B": B.tolist(), "mmd2": f(v), "str": s(v)})
    for G in (1, 5, 64):
        lw = rng.normal(0, 0.3, G); a = rng.normal(-2, 3, G); t = a + rng.normal(0, 1.0, G)
        v = mpo.calc_tv(list(lw), list(a), list(t))
        out["tv"].append({"log_ws": lw.tolist(), "a": a.tolist(), "t": t.tolist(), "tv": f(v), "str": s(v)})
    return out


if __name__ == "__main__":
    for name, fn in (("likelihood_fixture", likelihood_fixture), ("gp_small", gp_small), ("metrics_small", metrics_small)):
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(fn(), fh, indent=1)
        print("wrote", name)
