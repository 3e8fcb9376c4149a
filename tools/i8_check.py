"""Developer check on a B200: INT8 variance engine vs the FP64 (DMMA) engine and vs the oracle, plus timings.
Prints JSON lines; not a test (tests/ holds those).  usage: python tools/i8_check.py [quick]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bosip_b200 as bb  # noqa: E402
from bosip_b200 import synth  # noqa: E402
from oracle import bosip_oracle as orc  # noqa: E402


def compare(name, M, ctx, slices_list=(0, 7, 8, 6), seed=3, near_train=True):
    pr = synth.make_problem(name)
    rng = np.random.default_rng(seed)
    Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, M))
    if near_train:  # a few queries on / next to training points: the cancellation alpha^2 - q is worst there
        n = min(64, pr.N, M)
        Xq[:, :n] = pr.X[:, :n]
        Xq[:, n:2 * n] = pr.X[:, :n] + 1e-6
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    omu, ovar = orc.gp_predict(ofit, Xq)
    ctx.set_variance_engine("fp64")
    mu0, var0 = post.mean_and_var(Xq)
    a2 = pr.alpha[:, None] ** 2
    tol_den = np.abs(ovar) + 1e-3 * a2
    out = {"config": name, "M": M, "N": pr.N, "p": pr.p,
           "fp64_var_err": float(np.max(np.abs(var0 - ovar) / tol_den)), "fp64_mu_err": float(np.max(np.abs(mu0 - omu)) / np.max(np.abs(omu)))}
    for S in slices_list:
        ctx.set_variance_engine("int8", S)
        mu1, var1 = post.mean_and_var(Xq)
        out[f"i8s{S}_var_err"] = float(np.max(np.abs(var1 - ovar) / tol_den))
        out[f"i8s{S}_vs_fp64"] = float(np.max(np.abs(var1 - var0) / tol_den))
        out[f"i8s{S}_mu_eq"] = bool(np.array_equal(mu1, mu0))
        out[f"i8s{S}_nan"] = int(np.isnan(var1).sum())
    ctx.set_variance_engine("fp64")
    print(json.dumps(out), flush=True)
    return pr, post


def timing(name, M, ctx, slices_list=(7, 8)):
    import torch
    pr = synth.make_problem(name)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    Xd = torch.empty((M, pr.d), dtype=torch.float64, device="cuda").t()
    bb.candidates(ctx, bb.CandidateSource.philox(99, pr.lb, pr.ub, M), 0, M, device_out=Xd)
    for eng, S in [("fp64", 0)] + [("int8", s) for s in slices_list]:
        ctx.set_variance_engine(eng, S)
        bb.posterior_eval(post, like, prior, Xd, 7)
        ctx.timing_reset(True)
        torch.cuda.synchronize(); t0 = time.time()
        bb.posterior_eval(post, like, prior, Xd, 7)
        torch.cuda.synchronize(); dt = time.time() - t0
        tm = ctx.timing_get(); ctx.timing_reset(False)
        print(json.dumps({"cfg": name, "engine": eng, "slices": S, "M": M, "pts_per_s": M / dt,
                          "ms": {k: v["ms"] for k, v in tm.items()}, "launches": {k: v["launches"] for k, v in tm.items()}}), flush=True)
    ctx.set_variance_engine("fp64")


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ctx = bb.Context(0)
    if len(sys.argv) > 1 and sys.argv[1] == "one":   # one comparison: one <cfg> <M> <slices>
        compare(sys.argv[2], int(sys.argv[3]), ctx, slices_list=(int(sys.argv[4]),))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        for spec in sys.argv[2:] or ["C3"]:
            cfg, _, m = spec.partition(":")
            timing(cfg, int(m) if m else {"C3": 148 * 128 * 8, "C2": 10 ** 6, "C5": 148 * 128 * 2}[cfg], ctx, slices_list=(0, 7))   # 0: default (7 x 6 slices)
        return
    print(json.dumps({"int8_peak_tops": ctx.int8_peak(), **ctx.fp64_peak()}), flush=True)
    compare("C1", 1000, ctx)
    compare("C2", 5000, ctx)
    compare("C3", 20000, ctx)
    if quick:
        return
    timing("C3", 148 * 128 * 8, ctx)
    timing("C2", 10 ** 6, ctx)
    compare("C5", 2000, ctx, slices_list=(0, 7))
    timing("C5", 148 * 128 * 2, ctx, slices_list=(7,))


if __name__ == "__main__":
    main()
