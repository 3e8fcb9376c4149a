"""Secondary measurements for the other BASELINE.json configs (C1, C2, C5, C4 MMD, TV) on one B200; JSON lines."""
import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bosip_b200 as bb

ctx = bb.Context(0)
peak = ctx.fp64_peak()
print(json.dumps({"fp64_peak": peak}), flush=True)


def gp_case(name, M, reps=3, **kw):
    pr = bb.synth.make_problem(name, **kw)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
    t0 = time.perf_counter(); post2 = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx); t_fit2 = time.perf_counter() - t0
    del post2
    Xd = torch.empty((M, pr.d), dtype=torch.float64, device="cuda").t()
    bb.candidates(ctx, bb.CandidateSource.philox(1, pr.lb, pr.ub, M), 0, M, device_out=Xd)
    out = {"config": name, "d": pr.d, "p": pr.p, "N": pr.N, "M": M, "fit_s_first": t_fit, "fit_s": t_fit2}
    for want, label in ((7, "posterior+maxvar"), (1, "approx_posterior_only")):
        bb.posterior_eval(post, like, prior, Xd, want, acq=bb.MaxVar() if want == 7 else None)
        ctx.timing_reset(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            bb.posterior_eval(post, like, prior, Xd, want, acq=bb.MaxVar() if want == 7 else None)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        tm = ctx.timing_get(); ctx.timing_reset(False)
        fl_var = pr.p * pr.N * (pr.N + 1)
        rec = {"pts_per_s": M / dt, "ms": dt * 1e3, "kernel_ms": {k: v["ms"] / reps for k, v in tm.items()}}
        if want == 7 and tm["var_gemm"]["ms"] > 0:
            rec["var_gemm_tflops_algorithmic"] = fl_var * M / (tm["var_gemm"]["ms"] / reps * 1e-3) * 1e-12
            rec["var_gemm_frac_of_dmma_peak"] = rec["var_gemm_tflops_algorithmic"] / peak["dmma_tflops"]
        out[label] = rec
    print(json.dumps(out), flush=True)


which = sys.argv[1:] or ["C1", "C2", "C3", "C5", "C5p4", "MMD", "TV", "AMIS"]
if "C1" in which: gp_case("C1", 10201, reps=20)
if "C2" in which: gp_case("C2", 10 ** 6)
if "C3" in which: gp_case("C3", 148 * 128 * 16)
if "C5" in which: gp_case("C5", 148 * 128 * 8)
if "C5p4" in which: gp_case("C5", 148 * 128 * 4, p=4)
if "MMD" in which:
    rng = np.random.default_rng(0)
    for d in (5, 2):
        n = m = 100_000
        A = torch.as_tensor(rng.standard_normal((n, d)), device="cuda").t()
        B = torch.as_tensor(0.1 + 1.1 * rng.standard_normal((m, d)), device="cuda").t()
        k = bb.Kernel(bb.SE, np.ones(d))
        v = bb.calc_mmd(k, A, B, ctx)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        v = bb.calc_mmd(k, A, B, ctx)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        evals = n * n + m * m + n * m
        print(json.dumps({"config": "C4 MMD", "d": d, "n": n, "m": m, "mmd2": v, "s": dt, "kernel_evals_per_s": evals / dt,
                          "flops_as_written_tflops": evals * (3 * d + 3) / dt * 1e-12}), flush=True)
if "TV" in which:
    rng = np.random.default_rng(1)
    for G in (10 ** 6, 10 ** 8):
        lw = torch.as_tensor(rng.normal(0, 0.1, G), device="cuda"); a = torch.as_tensor(rng.normal(-3, 2, G), device="cuda")
        t = a + 0.5 * torch.randn(G, dtype=torch.float64, device="cuda")
        v = bb.calc_tv(lw, a, t, ctx)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        v = bb.calc_tv(lw, a, t, ctx)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"config": "TV", "G": G, "tv": v, "s": dt, "algorithmic_GBps": 24 * G / dt * 1e-9, "frac_of_hbm_6451": 24 * G / dt * 1e-9 / 6451.2}), flush=True)
if "AMIS" in which:
    # AMIS on the C3 posterior mean: T = 10 iterations x N = 1e5 draws, everything resident in HBM
    pr = bb.synth.make_problem("C3")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    problem = bb.BosipProblem(pr.X, pr.Y, model, bb.NormalLikelihood(pr.z_obs, pr.std_obs), bb.ProductPrior(pr.prior_kind, pr.prior_params),
                              bounds=(pr.lb, pr.ub), ctx=ctx)
    log_pi = bb.log_posterior_mean(problem)
    q0 = bb.NormalProposal(0.5 * (pr.lb + pr.ub), np.diag((pr.ub - pr.lb) / 5.0))
    amis = bb.AMIS(T=10, N=100_000)
    amis(log_pi, q0, bb.AnalyticalFitter(), seed=1, ctx=ctx)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xs, ws = amis(log_pi, q0, bb.AnalyticalFitter(), seed=2, ctx=ctx)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # share of log_pi (the fused posterior sweep) in it
    t0 = time.perf_counter()
    for _ in range(11):
        log_pi(xs[:, :100_000])
    torch.cuda.synchronize(); dt_pi = time.perf_counter() - t0
    print(json.dumps({"config": "AMIS on C3 log_posterior_mean", "T": 10, "N": 100_000, "s": dt, "draws_per_s": 11 * 100_000 / dt,
                      "log_pi_share": dt_pi / dt, "ess": float(1.0 / (ws ** 2).sum().item())}), flush=True)
