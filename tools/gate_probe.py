"""Error distributions behind the parity gates (tests/test_gpu_parity.py) + fit timings: prints JSON lines.
   python tools/gate_probe.py            (on a B200)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bosip_b200 as bb
from oracle import bosip_oracle as orc

ctx = bb.Context(0)


def case(name, M, engine, **kw):
    pr = bb.synth.make_problem(name, **kw)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    olike = orc.NormalLikelihood(pr.z_obs, pr.std_obs); oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    rng = np.random.default_rng(7)
    w = (pr.ub - pr.lb)[:, None]
    Xq = rng.uniform(pr.lb[:, None] - 0.05 * w, pr.ub[:, None] + 0.05 * w, size=(pr.d, M))
    Xq[:, :min(M, pr.N) // 4] = pr.X[:, :min(M, pr.N) // 4]            # queries ON training points: the worst cancellation
    ctx.set_variance_engine(engine)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    Lg, Lig, ag = post.factors()
    mu, var = post.mean_and_var(Xq)
    omu, ovar = orc.gp_predict(ofit, Xq)
    kxx = (pr.alpha ** 2)[:, None]
    big = ovar >= 1e-6 * kxx
    res = bb.posterior_eval(post, like, prior, Xq, 7)
    ores = orc.posterior_eval(ofit, olike, oprior, Xq)
    a, b = res["log_post_var"], ores["log_post_var"]
    fin = np.isfinite(a) & np.isfinite(b)
    std = np.sqrt(np.maximum(ores["var"], 0.0))
    lm = orc.log_likelihood_mean(olike, ores["mu"], std); lsq = orc.log_sq_likelihood_mean(olike, ores["mu"], std)
    with np.errstate(all="ignore"):
        amp = (np.exp(lsq) + np.exp(2 * lm)) / (np.exp(lsq) - np.exp(2 * lm))
    err = np.abs(a - b)
    out = {"case": name, "kw": kw, "engine": engine, "N": pr.N, "M": M,
           "a_relerr_max": float(np.max(np.abs(ag - ofit.a)) / np.max(np.abs(ofit.a))),
           "L_relerr_max": float(np.max(np.abs(Lg - ofit.L)) / np.max(np.abs(ofit.L))),
           "mu_err_gate": float(np.max(np.abs(mu - omu) / (np.abs(omu) + 1e-3 * np.max(np.abs(omu))))),
           "var_min_over_kxx": float(np.min(ovar / kxx)),
           "var_relerr_where_ge_1e-6": float(np.max((np.abs(var - ovar) / np.abs(ovar))[big])) if big.any() else None,
           "var_abserr_over_kxx_elsewhere": float(np.max((np.abs(var - ovar) / kxx)[~big])) if (~big).any() else None,
           "var_abserr_over_kxx_all": float(np.max(np.abs(var - ovar) / kxx)),
           "old_gate_units": float(np.max(np.abs(var - ovar) / (np.abs(ovar) + 1e-3 * kxx))),
           "logvar_relerr_max": float(np.max(err[fin] / np.abs(b[fin]))), "logvar_abserr_max": float(np.max(err[fin])),
           "logvar_abserr_over_amp_max": float(np.max(err[fin] / amp[fin])), "amp_max": float(np.max(amp[fin])),
           "finite_mismatch": int(np.sum(np.isfinite(a) != np.isfinite(b)))}
    print(json.dumps(out), flush=True)
    post.free()


for engine in ("fp64", "int8"):
    case("C1", 10201, engine)
    case("C2", 30000, engine)
    case("C3", 20000, engine)
    case("C2", 20000, engine, noise_frac=1e-3)
    case("X", 4000, engine, d=6, p=1, N=130, kernel_id=0, seed=1000 + 31 * 6 + 130)
    case("X", 4000, engine, d=7, p=2, N=257, kernel_id=1, seed=1000 + 31 * 7 + 257)
ctx.set_variance_engine("auto")
# ---- fit timings (host wall clock around the synchronous call, second call of each shape)
for (N, p, d) in ((1000, 4, 5), (2048, 1, 10), (2049, 1, 10), (4096, 1, 10), (4096, 4, 10)):
    pr = bb.synth.make_problem("X", d=d, p=p, N=N, seed=5)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx); post.free()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx); ts.append(time.perf_counter() - t0); post.free()
    print(json.dumps({"fit": {"N": N, "p": p, "d": d}, "ms_best": 1e3 * min(ts), "ms_all": [1e3 * t for t in ts]}), flush=True)
