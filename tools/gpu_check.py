"""Developer check on a B200: CUDA path vs oracle on the synthetic configs + quick per-kernel timings.
Not a test (tests/ holds those); prints JSON lines."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bosip_b200 as bb  # noqa: E402
from bosip_b200 import synth  # noqa: E402
from oracle import bosip_oracle as orc  # noqa: E402


def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    fin = np.isfinite(a) & np.isfinite(b)
    same_inf = (~fin) & ((a == b) | (np.isnan(a) & np.isnan(b)))
    bad = (~fin) & (~same_inf)
    return float(np.max(np.abs(a - b)[fin] / den[fin])) if fin.any() else 0.0, int(bad.sum())


def check(name, M, ctx, seed=1):
    pr = synth.make_problem(name)
    rng = np.random.default_rng(seed)
    Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, M))
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    t0 = time.time()
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    t_fit = time.time() - t0
    t0 = time.time()
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    t_ofit = time.time() - t0
    Lg, Lig, ag = post.factors()
    out = {"config": name, "M": M, "N": pr.N, "fit_s": round(t_fit, 4), "oracle_fit_s": round(t_ofit, 4)}
    out["L_relerr"] = float(np.max(np.abs(Lg - ofit.L)) / np.max(np.abs(ofit.L)))
    out["a_relerr"] = float(np.max(np.abs(ag - ofit.a)) / np.max(np.abs(ofit.a)))
    # Linv check: ||Linv L - I||
    out["LinvL_minus_I"] = float(max(np.max(np.abs(Lig[j] @ ofit.L[j] - np.eye(pr.N))) for j in range(pr.p)))
    mu, var = post.mean_and_var(Xq)
    omu, ovar = orc.gp_predict(ofit, Xq)
    out["mu_relerr"] = relerr(mu, omu)[0]
    out["mu_abs_err_over_scale"] = float(np.max(np.abs(mu - omu)) / np.max(np.abs(omu)))
    out["var_abs_err_over_kxx"] = float(np.max(np.abs(var - ovar) / (pr.alpha[:, None] ** 2)))
    out["var_relerr"] = relerr(var, ovar)[0]
    res = bb.posterior_eval(post, like, prior, Xq, 7)
    oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    olike = orc.NormalLikelihood(pr.z_obs, pr.std_obs)
    ores = orc.posterior_eval(ofit, olike, oprior, Xq)
    for k in ("log_approx_post", "log_post_mean", "log_post_var"):
        e, bad = relerr(res[k], ores[k])
        out[k + "_relerr"] = e; out[k + "_inf_mismatch"] = bad
    out["n_clamped"] = res["n_clamped"]
    out["n_neg_inf_var"] = int(np.isneginf(ores["log_post_var"]).sum())
    # exact-parity mode: host factors
    post2 = bb.model_posterior_from_factors(model, pr.X, ofit.L, ofit.a, ctx=ctx)
    mu2, var2 = post2.mean_and_var(Xq)
    out["hostfac_mu_relerr"] = relerr(mu2, omu)[0]
    out["hostfac_var_abs_over_kxx"] = float(np.max(np.abs(var2 - ovar) / (pr.alpha[:, None] ** 2)))
    # argmax
    src = bb.CandidateSource.from_points(Xq)
    bi, bv, bx = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src)
    oi, ov = orc.maxvar_argmax(ores["log_post_var"], True)
    out["argmax_gpu"] = [bi, bv]; out["argmax_oracle"] = [oi, ov]
    print(json.dumps(out), flush=True)
    return pr, post, like, prior


def main():
    ctx = bb.Context(0)
    print(json.dumps(ctx.fp64_peak()), flush=True)
    check("C1", 10201, ctx)
    check("C2", 20000, ctx)
    pr, post, like, prior = check("C3", 20000, ctx)

    # philox candidates parity
    src = bb.CandidateSource.philox(1234, pr.lb, pr.ub, 10 ** 6)
    cg = bb.candidates(ctx, src, 5000, 1000)
    ch = synth.philox_uniform_candidates(1234, 5000, 1000, pr.lb, pr.ub)
    print(json.dumps({"philox_max_abs_diff": float(np.max(np.abs(cg - ch)))}), flush=True)

    # MMD + TV
    rng = np.random.default_rng(7)
    A = rng.standard_normal((5, 3000)); B = 0.1 + 1.1 * rng.standard_normal((5, 2500))
    k = bb.Kernel(bb.SE, np.ones(5))
    g = bb.calc_mmd(k, A, B, ctx)
    o = orc.calc_mmd(orc.SE, np.ones(5), A, B)[0]
    print(json.dumps({"mmd_gpu": g, "mmd_oracle": o, "rel": abs(g - o) / abs(o)}), flush=True)
    G = 200000
    lw = rng.normal(0, 0.1, G); a = rng.normal(-3, 2, G); t = a + rng.normal(0, 0.5, G)
    g = bb.calc_tv(lw, a, t, ctx); o = orc.calc_tv(lw, a, t)
    print(json.dumps({"tv_gpu": g, "tv_oracle": o, "rel": abs(g - o) / abs(o)}), flush=True)

    # throughput, C3, device-resident inputs
    import torch
    M = 148 * 128 * 8
    Xd = torch.empty((M, 5), dtype=torch.float64, device="cuda").t()
    bb.candidates(ctx, bb.CandidateSource.philox(99, pr.lb, pr.ub, M), 0, M, device_out=Xd)
    for want, label in ((7, "approx+mean+var"), (1, "approx only")):
        bb.posterior_eval(post, like, prior, Xd, want)
        ctx.timing_reset(True)
        torch.cuda.synchronize(); t0 = time.time()
        bb.posterior_eval(post, like, prior, Xd, want)
        torch.cuda.synchronize(); dt = time.time() - t0
        tm = ctx.timing_get(); ctx.timing_reset(False)
        print(json.dumps({"throughput": label, "M": M, "s": dt, "pts_per_s": M / dt, "kernels": tm}), flush=True)


if __name__ == "__main__":
    main()
