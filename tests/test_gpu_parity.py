"""GPU: the CUDA path through the C ABI vs the oracle on seeded inputs, the mpmath goldens, and the edge cases
(ragged sizes, chunk boundaries, every kernel/prior kind, options, device-resident buffers, error behaviour)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10  # BASELINE.json north_star: 1e-10 relative error in Float64


def problem(bb, name="C2", **kw):
    pr = bb.synth.make_problem(name, **kw)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    olike = orc.NormalLikelihood(pr.z_obs, pr.std_obs)
    oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    return pr, model, like, prior, ofit, olike, oprior


def queries(pr, M, seed=0, spill=0.2):
    """Uniform in a box slightly larger than the domain, so that some points fall outside the prior support."""
    rng = np.random.default_rng(seed)
    w = (pr.ub - pr.lb)[:, None]
    return rng.uniform(pr.lb[:, None] - spill * w, pr.ub[:, None] + spill * w, size=(pr.d, M))


def assert_mu(mu, omu):
    scale = np.max(np.abs(omu))
    assert np.max(np.abs(mu - omu) / (np.abs(omu) + 1e-3 * scale)) < TOL


def assert_var(var, ovar, kxx):
    # 1e-10 RELATIVE wherever sigma^2 / k(x,x) >= 1e-6.  Below that the subtraction alpha^2 - ||v||^2 has
    # cancelled six digits and both sides carry kappa(L) * eps of it: the bound there is absolute, 1e-10 of the 1e-6 k(x,x) floor
    # times the 1e3 the old gate allowed (never reached by these problems: their smallest variance is ~6e-4 k(x,x), on top of the
    # training points).  Observed with queries ON training points (tools/gate_probe.py, profiles/r02_gate_probe.jsonl):
    # <= 1.5e-11 relative on either engine.
    k = kxx[:, None] * np.ones_like(ovar)
    big = ovar >= 1e-6 * k
    assert np.max((np.abs(var - ovar) / np.maximum(np.abs(ovar), 1e-300))[big], initial=0.0) < TOL
    assert np.max((np.abs(var - ovar) / k)[~big], initial=0.0) < 1e-13


def assert_logs(res, ores):
    for k in ("log_approx_post", "log_post_mean"):
        a, b = res[k], ores[k]
        fin = np.isfinite(b)
        np.testing.assert_array_equal(np.isneginf(a), np.isneginf(b))
        np.testing.assert_allclose(a[fin], b[fin], rtol=TOL)
    a, b = res["log_post_var"], ores["log_post_var"]
    # log V = log(E[l^2] - E[l]^2) is computed in linear space like the reference; errors are amplified by
    # amp = (E[l^2] + E[l]^2)/V, so the gate is 1e-10 relative + eps-level * amp
    mu, var = ores["mu"], ores["var"]
    fin = np.isfinite(b) & np.isfinite(a)
    assert np.sum(np.isfinite(a) != np.isfinite(b)) <= 1e-3 * a.size + 1  # V rounds to 0 on one side only at the underflow edge
    err = np.abs(a[fin] - b[fin])
    assert np.all(err <= TOL * np.abs(b[fin]) + 1e-7)
    if err.size:
        assert np.median(err / np.abs(b[fin])) < 1e-12


@pytest.mark.parametrize("name,M", [("C1", 10201), ("C2", 30000), ("C3", 20000)])
def test_fit_predict_and_closures_vs_oracle(bb, ctx, name, M):
    pr, model, like, prior, ofit, olike, oprior = problem(bb, name)
    Xq = bb.synth.grid_queries(pr.lb, pr.ub, 101) if name == "C1" else queries(pr, M)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    Lg, Lig, ag = post.factors()
    assert np.max(np.abs(Lg - ofit.L)) / np.max(np.abs(ofit.L)) < 1e-12
    assert np.max(np.abs(ag - ofit.a)) / np.max(np.abs(ofit.a)) < 1e-11  # two triangular solves with L (observed <= 3e-13); was 1e-9 through L^-1
    for j in range(pr.p):
        assert np.max(np.abs(Lig[j] @ ofit.L[j] - np.eye(pr.N))) < 1e-11
    omu, ovar = orc.gp_predict(ofit, Xq)
    mu, var = post.mean_and_var(Xq)
    assert mu.shape == (pr.p, Xq.shape[1]) and var.shape == mu.shape          # p x M like test/unit/utils.jl:8-11
    assert_mu(mu, omu); assert_var(var, ovar, pr.alpha ** 2)
    np.testing.assert_allclose(post.mean(Xq), mu, rtol=0, atol=0)              # mean-only path is bit-identical
    np.testing.assert_allclose(post.mean_and_std(Xq)[1], np.sqrt(var), rtol=1e-15)
    v1 = post.mean_and_var(Xq[:, 7])                                            # vector method -> length p
    assert v1[0].shape == (pr.p,) and np.array_equal(v1[0], mu[:, 7]) and np.array_equal(v1[1], var[:, 7])
    res = bb.posterior_eval(post, like, prior, Xq, 7)
    ores = orc.posterior_eval(ofit, olike, oprior, Xq)
    assert_logs(res, ores)
    assert np.isneginf(res["log_approx_post"]).sum() == np.isneginf(ores["log_prior"]).sum()


@pytest.mark.parametrize("name,M", [("C2", 5000), ("C3", 5000)])
def test_host_factors_exact_parity_mode(bb, ctx, name, M):
    """Evaluator parity with the SAME factors as the oracle (bosip_b200_gp_load_factors): kappa-free 1e-10 gate."""
    pr, model, like, prior, ofit, olike, oprior = problem(bb, name)
    Xq = queries(pr, M, seed=3)
    post = bb.model_posterior_from_factors(model, pr.X, ofit.L, ofit.a, ctx=ctx)
    mu, var = post.mean_and_var(Xq)
    omu, ovar = orc.gp_predict(ofit, Xq)
    assert_mu(mu, omu); assert_var(var, ovar, pr.alpha ** 2)
    assert_logs(bb.posterior_eval(post, like, prior, Xq, 7), orc.posterior_eval(ofit, olike, oprior, Xq))


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_mpmath_goldens(bb, ctx, ci):
    c = json.load(open(os.path.join(GOLD, "gp_small.json")))["cases"][ci]
    X, Y, Xq = np.array(c["X"]), np.array(c["Y"]), np.array(c["Xq"])
    model = bb.GaussianProcess(c["kernel_id"], np.array(c["lam"]), np.array(c["alpha"]), np.array(c["sigma"]), clamp_var=False)
    post = bb.model_posterior(model, X, Y, ctx=ctx)
    like = bb.NormalLikelihood(c["z_obs"], c["std_obs"])
    mu, var = post.mean_and_var(Xq)
    a2 = np.array(c["alpha"]) ** 2
    for name, prd in c["priors"].items():
        prior = bb.ProductPrior(prd["kind"], prd["params"])
        res = bb.posterior_eval(post, like, prior, Xq, 7)
        for i, pt in enumerate(c["points"]):
            np.testing.assert_allclose(mu[:, i], pt["mu"], rtol=TOL, atol=1e-12)
            np.testing.assert_allclose(var[:, i], pt["var"], rtol=TOL, atol=1e-12 * a2.max())
            lp = pt["log_prior"][name]
            if math.isinf(lp):
                assert res["log_approx_post"][i] == -np.inf and res["log_post_mean"][i] == -np.inf and res["log_post_var"][i] == -np.inf
                continue
            assert res["log_approx_post"][i] == pytest.approx(lp + pt["loglike"], rel=TOL)
            assert res["log_post_mean"][i] == pytest.approx(lp + pt["log_like_mean"], rel=TOL)
            amp = math.exp(pt["log_sq_like_mean"]) / max(pt["like_var"], 1e-300)
            assert res["log_post_var"][i] == pytest.approx(2 * lp + pt["log_like_var"], abs=1e-11 * amp + 1e-9)


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
@pytest.mark.parametrize("d,p,N,M", [(1, 1, 1, 1), (3, 2, 63, 127), (3, 2, 64, 128), (4, 3, 65, 129), (6, 1, 130, 1000),
                                      (7, 2, 257, 513), (12, 1, 100, 300), (20, 1, 48, 77)])
def test_ragged_shapes_and_kernels(bb, ctx, kernel_id, d, p, N, M):
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "X", d=d, p=p, N=N, kernel_id=kernel_id, seed=1000 + 31 * d + N)
    Xq = queries(pr, M, seed=N, spill=0.2 if d <= 4 else 0.01)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    mu, var = post.mean_and_var(Xq)
    omu, ovar = orc.gp_predict(ofit, Xq)
    assert_mu(mu, omu); assert_var(var, ovar, pr.alpha ** 2)
    assert_logs(bb.posterior_eval(post, like, prior, Xq, 7), orc.posterior_eval(ofit, olike, oprior, Xq))


def test_chunking_is_invisible(bb, ctx):
    """Results are bit-identical whatever the internal chunk size, and for host vs device-resident buffers."""
    import torch
    pr, model, like, prior, *_ = problem(bb, "C2")
    Xq = queries(pr, 3001, seed=9)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    ref = bb.posterior_eval(post, like, prior, Xq, 7)
    try:
        for chunk in (128, 256, 1000):
            ctx.set_chunk(chunk)
            r = bb.posterior_eval(post, like, prior, Xq, 7)
            for k in ("log_approx_post", "log_post_mean", "log_post_var"):
                np.testing.assert_array_equal(r[k], ref[k])
            mu, var = post.mean_and_var(Xq)
            np.testing.assert_array_equal(mu, post.mean(Xq))
    finally:
        ctx.set_chunk(0)
    Xd = torch.as_tensor(np.ascontiguousarray(Xq.T), device="cuda").t()
    rd = bb.posterior_eval(post, like, prior, Xd, 7)
    for k in ("log_approx_post", "log_post_mean", "log_post_var"):
        assert rd[k].is_cuda
        np.testing.assert_array_equal(rd[k].cpu().numpy(), ref[k])
    mud, vard = post.mean_and_var(Xd)
    mu, var = post.mean_and_var(Xq)
    np.testing.assert_array_equal(mud.cpu().numpy(), mu); np.testing.assert_array_equal(vard.cpu().numpy(), var)
    # want-subsets agree with the full call
    for want, key in ((1, "log_approx_post"), (2, "log_post_mean"), (4, "log_post_var")):
        np.testing.assert_array_equal(bb.posterior_eval(post, like, prior, Xq, want)[key], ref[key])


def test_options_priors_and_mean_function(bb, ctx):
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "C2")
    Xq = queries(pr, 2000, seed=4)
    # pred_noise + jitter + no clamp
    m2 = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma, jitter=1e-6, pred_noise=True, clamp_var=False)
    o2 = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma, opts=orc.GPOpts(1e-6, True, False))
    mu, var = bb.model_posterior(m2, pr.X, pr.Y, ctx=ctx).mean_and_var(Xq)
    omu, ovar = orc.gp_predict(o2, Xq)
    assert_mu(mu, omu); assert_var(var, ovar, pr.alpha ** 2)
    # prior mean function m_j(x)
    mean_fn = lambda X: np.stack([0.3 * X[0] - 0.1, np.cos(X[1])])
    m3 = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma, mean=mean_fn)
    o3 = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma, mean_at_X=mean_fn(pr.X))
    mu3 = bb.model_posterior(m3, pr.X, pr.Y, ctx=ctx).mean(Xq)
    assert_mu(mu3, orc.gp_predict(o3, Xq, mean_at_Xq=mean_fn(Xq), want_var=False))
    # uniform / normal / mixed priors and host-supplied log-prior values
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    for kinds, prms in (([0, 0], [[-5, 5], [-4, 6]]), ([1, 1], [[0.5, 2.0], [-1.0, 3.0]]), ([2, 0], [[0, 2, -5, 5], [-5, 5]])):
        a = bb.posterior_eval(post, like, bb.ProductPrior(kinds, prms), Xq, 7)
        b = orc.posterior_eval(ofit, olike, orc.ProductPrior(kinds, prms), Xq)
        assert_logs(a, b)
        c = bb.posterior_eval(post, like, None, Xq, 7, logprior=b["log_prior"])
        assert_logs(c, b)
    # no prior at all == the likelihood-level closures
    nl = bb.posterior_eval(post, like, None, Xq, 7)
    np.testing.assert_array_equal(bb.log_likelihood_variance(like, post)(Xq), nl["log_post_var"])
    np.testing.assert_array_equal(bb.log_approx_likelihood(like, post)(Xq), nl["log_approx_post"])


def test_closures_normalisation_and_evidence(bb, ctx):
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "C2")
    prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx)
    Xq = queries(pr, 500, seed=2, spill=0.0)
    ores = orc.posterior_eval(ofit, olike, oprior, Xq)
    np.testing.assert_allclose(bb.approx_posterior(prob)(Xq), np.exp(ores["log_approx_post"]), rtol=1e-9)
    np.testing.assert_allclose(bb.posterior_mean(prob)(Xq), np.exp(ores["log_post_mean"]), rtol=1e-9)
    np.testing.assert_allclose(bb.posterior_variance(prob)(Xq), np.exp(ores["log_post_var"]), rtol=1e-6)
    assert isinstance(bb.log_posterior_mean(prob)(Xq[:, 3]), float)            # vector method -> Real
    xs = prior.rand(10_000, np.random.default_rng(1))                           # evidence.jl:21
    oxs = orc.posterior_eval(ofit, olike, oprior, xs)
    for which, key in (("approx", "log_approx_post"), ("mean", "log_post_mean")):
        ev = bb.evidence(prob, which, xs=xs)
        assert ev == pytest.approx(orc.evidence(oxs[key], oxs["log_prior"]), rel=1e-10)
    ev = bb.evidence(prob, "mean", xs=xs)
    np.testing.assert_allclose(bb.posterior_mean(prob, normalize=True, xs=xs)(Xq), np.exp(ores["log_post_mean"] - math.log(ev)), rtol=1e-9)
    np.testing.assert_allclose(bb.posterior_variance(prob, normalize=True, xs=xs)(Xq), np.exp(ores["log_post_var"] - 2 * math.log(ev)), rtol=1e-6)


def test_argmax_matches_oracle_ties_and_sources(bb, ctx):
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "C2")
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    Xq = queries(pr, 40000, seed=11, spill=0.05)
    ores = orc.posterior_eval(ofit, olike, oprior, Xq)
    top2 = np.sort(ores["log_post_var"][np.isfinite(ores["log_post_var"])])[-2:]
    assert top2[1] - top2[0] > 1e-9 * abs(top2[1]), "regenerate: not tie-free"
    for acq, logspace in ((bb.LogMaxVar(), True), (bb.MaxVar(), False)):
        oi, ov = orc.maxvar_argmax(ores["log_post_var"], logspace)
        bi, bv, bx = bb.acq_argmax(post, like, prior, acq, bb.CandidateSource.from_points(Xq))
        assert bi == oi and bv == pytest.approx(ov, rel=1e-9)
        np.testing.assert_array_equal(bx, Xq[:, oi])
    # duplicated winner -> first index (Julia argmax); offset labels global indices
    Xt = np.concatenate([Xq[:, :100], Xq[:, [oi]], Xq[:, 100:200], Xq[:, [oi]]], axis=1)
    bi, _, _ = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(Xt))
    assert bi == 100
    # every candidate outside the prior support: all values are -Inf and Julia's argmax returns the FIRST index (labelled by
    # index_offset); an empty candidate set is the only case without a winner
    far = np.full((2, 300), 9.0)
    bi, bv, bx = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(far))
    assert bi == 0 and bv == -np.inf and np.array_equal(bx, far[:, 0])
    bi, bv, _ = bb.acq_argmax(post, like, prior, bb.MaxVar(), bb.CandidateSource.from_points(far), 100, 200)
    assert bi == 100 and bv == 0.0                                           # MaxVar = exp(log V) = 0 everywhere
    mixed = np.concatenate([far[:, :5], Xq[:, :50]], axis=1)
    assert bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(mixed))[0] >= 5
    bi, bv, _ = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(far[:, :0]))
    assert bi == -1 and bv == -np.inf
    # device-generated candidates == host Philox mirror == explicit points, and sharded ranges compose
    src = bb.CandidateSource.philox(20261019, pr.lb, pr.ub, 50000)
    cand = bb.synth.philox_uniform_candidates(20261019, 0, 50000, pr.lb, pr.ub)
    np.testing.assert_array_equal(bb.candidates(ctx, src, 0, 50000), cand)
    full = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src)
    pts = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(cand))
    assert full[0] == pts[0] and full[1] == pts[1]
    np.testing.assert_array_equal(full[2], cand[:, full[0]])
    shards = [bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src, *bb.parallel.shard_range(50000, r, 8)) for r in range(8)]
    w = bb.parallel.combine_argmax(np.array([s[1] for s in shards]), np.array([s[0] for s in shards]))
    assert shards[w][0] == full[0] and shards[w][1] == full[1]
    # grid source (C1: 101 x 101, examples/simple/main.jl:69) == explicit grid
    gsrc = bb.CandidateSource.grid(pr.lb, pr.ub, 101)
    G = bb.synth.grid_queries(pr.lb, pr.ub, 101)
    np.testing.assert_array_equal(bb.candidates(ctx, gsrc, 0, 10201), G)
    a = bb.acq_argmax(post, like, prior, bb.MaxVar(), gsrc); b = bb.acq_argmax(post, like, prior, bb.MaxVar(), bb.CandidateSource.from_points(G))
    assert a[0] == b[0] and a[1] == b[1]
    # the maximiser object (single process: identity exchange)
    prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx)
    x, val = bb.B200CandidateAM(src).maximize_acquisition(prob, bb.LogMaxVar())
    np.testing.assert_array_equal(x, full[2]); assert val == full[1]


def test_error_behaviour(bb, ctx):
    pr, model, like, prior, *_ = problem(bb, "C1")
    # duplicate training points with zero noise: not positive definite -> PosDefException (Julia cholesky)
    Xd = np.concatenate([pr.X, pr.X[:, :1]], axis=1); Yd = np.concatenate([pr.Y, pr.Y[:, :1]], axis=1)
    with pytest.raises(bb.PosDefException):
        bb.model_posterior(bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, 0.0 * pr.sigma), Xd, Yd, ctx=ctx)
    with pytest.raises(bb.BosipB200Error) as e:
        bb.model_posterior(bb.GaussianProcess(7, pr.lam, pr.alpha, pr.sigma), pr.X, pr.Y, ctx=ctx)
    assert e.value.code == bb._lib.EINVAL
    with pytest.raises(bb.BosipB200Error):
        bb.model_posterior(bb.GaussianProcess(pr.kernel_id, -pr.lam, pr.alpha, pr.sigma), pr.X, pr.Y, ctx=ctx)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    with pytest.raises(bb.BosipB200Error):  # p mismatch between likelihood and GP
        bb.posterior_eval(post, bb.NormalLikelihood([0.0, 1.0], [1.0, 1.0]), prior, pr.X, 7)
    with pytest.raises(bb.BosipB200Error):  # std_obs must be positive
        bb.posterior_eval(post, bb.NormalLikelihood([0.0], [0.0]), prior, pr.X, 7)
    with pytest.raises(ValueError):         # wrong input dimension
        post.mean(np.zeros((3, 4)))
    # the context stays usable after errors, and an empty batch is a no-op
    assert post.mean(pr.X).shape == (1, pr.N)
    assert post.mean(np.zeros((2, 0))).shape == (1, 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_fast_math_accuracy(bb, ctx, variant):
    """csrc/fastmath.cuh (branch-free exp(-x), sqrt(x); variant 1 = the cross-kernel's table-driven exp and one-step sqrt) vs the
    CUDA library and vs NumPy: <= 2 ulp (variant 1: plus an absolute 1.3e-17 from its one-constant reduction), exact at the edges."""
    rng = np.random.default_rng(0)
    x = np.concatenate([[0.0, 1e-300, 1e-200, 1e-30, 1e-16, 0.5, 1.0, 2.0, 707.9, 708.0, 708.1, 745.0, 1e3, 1e6],
                        rng.uniform(0, 50, 200000), 10.0 ** rng.uniform(-12, 2.8, 200000)])
    n = x.size
    outs = [np.empty(n) for _ in range(4)]
    dp = bb._lib.c_double_p
    ctx._check(ctx.lib.bosip_b200_math_selftest(ctx.h, x.ctypes.data_as(dp), n, variant, *[o.ctypes.data_as(dp) for o in outs]))
    ef, qf, er, qr = outs
    big = x <= 700.0
    # variant 1 reduces the argument with ln2/64 as one double (3.6e-19 off): +3.4e-17 * x relative, i.e. <= 1.3e-17 absolute
    tol = 4.5e-16 + (3.4e-17 * x[big] if variant == 1 else 0.0)
    assert np.all(np.abs(ef[big] - er[big]) / er[big] < tol)
    assert np.all(np.abs(ef[big] - np.exp(-x[big])) / np.exp(-x[big]) < tol)
    assert np.max(np.abs(ef[big] - er[big])) < 2.3e-16
    assert np.all(ef[x > 708.1] < 1e-300 if variant == 1 else ef[x > 708.0] == 0.0) and ef[0] == 1.0   # variant 1 leaves a denormal
    ok = x >= 1e-280
    assert np.max(np.abs(qf[ok] - qr[ok]) / qr[ok]) < 2.3e-16
    assert (qf[0] == 0.0 if variant == 0 else qf[0] < 1e-140) and np.all(qf[~ok] < 1e-140)   # variant 1: sqrt(x + 1e-300)


def test_bi_hyperparameter_samples(bb, ctx):
    """MultiFittedParams (BI): log-mean-exp over S hyper-parameter samples in ONE fused call (a15 of SURVEY.md section 8a)."""
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "C2")
    rng = np.random.default_rng(21)
    S = 5
    models, ofits = [], []
    for s_ in range(S):
        lam = pr.lam * rng.uniform(0.7, 1.4, pr.lam.shape); alpha = pr.alpha * rng.uniform(0.8, 1.2, pr.p); sigma = pr.sigma * rng.uniform(0.5, 2.0, pr.p)
        models.append(bb.GaussianProcess(pr.kernel_id, lam, alpha, sigma))
        ofits.append(orc.gp_fit(pr.kernel_id, pr.X, pr.Y, lam, alpha, sigma))
    Xq = queries(pr, 5000, seed=5, spill=0.05)
    prob = bb.BosipProblem(pr.X, pr.Y, models, like, prior, (pr.lb, pr.ub), ctx)
    posts = prob.model_posterior()
    assert isinstance(posts, list) and len(posts) == S
    res = bb.posterior_eval(posts, like, prior, Xq, 7, acq=bb.LogMaxVar())
    ores = orc.posterior_eval_bi(ofits, olike, oprior, Xq)
    for k in ("log_approx_post", "log_post_mean"):
        fin = np.isfinite(ores[k])
        np.testing.assert_array_equal(np.isfinite(res[k]), fin)
        np.testing.assert_allclose(res[k][fin], ores[k][fin], rtol=TOL)
    fin = np.isfinite(ores["log_post_var"]) & np.isfinite(res["log_post_var"])
    assert fin.sum() > 0.5 * fin.size
    assert np.all(np.abs(res["log_post_var"][fin] - ores["log_post_var"][fin]) <= TOL * np.abs(ores["log_post_var"][fin]) + 1e-7)
    oi, ov = orc.maxvar_argmax(ores["log_post_var"], True)
    assert res["best_idx"] == oi
    # the closures of the problem object take the ensemble path, and a one-member ensemble equals the MAP path
    np.testing.assert_array_equal(bb.log_posterior_variance(prob)(Xq), res["log_post_var"])
    one = bb.posterior_eval([posts[0]], like, prior, Xq, 7)
    ref = bb.posterior_eval(posts[0], like, prior, Xq, 7)
    for k in ("log_approx_post", "log_post_mean", "log_post_var"):
        np.testing.assert_allclose(one[k], ref[k], rtol=1e-13)
    # chunking stays invisible
    try:
        ctx.set_chunk(512)
        r2 = bb.posterior_eval(posts, like, prior, Xq, 7)
        for k in ("log_approx_post", "log_post_mean", "log_post_var"):
            np.testing.assert_array_equal(r2[k], res[k])
    finally:
        ctx.set_chunk(0)


@pytest.mark.parametrize("func,key", [("posterior_mean", "log_post_mean"), ("approx_posterior", "log_approx_post"), ("posterior_variance", "log_post_var")])
def test_marginal_integration_sweep(bb, ctx, func, key):
    """compute_marginals_int (ext/CairoExt.jl:146-246): device-generated sweep == the reference's overwrite-rows loop."""
    pr, model, like, prior, ofit, olike, oprior = problem(bb, "X", d=3, p=2, N=90, seed=77)
    prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx)
    lhc = bb.generate_LHC((pr.lb, pr.ub), 37, np.random.default_rng(3))       # G = 37: not a divisor of anything
    gs = 9
    res = bb.compute_marginals_int(prob, func=func, grid_size=gs, xs=lhc, normalize=False)
    f = lambda X: np.exp(orc.posterior_eval(ofit, olike, oprior, X)[key])
    od, op = orc.marginals_int(f, lhc, pr.lb, pr.ub, gs)
    tol = TOL if key != "log_post_var" else 1e-6
    for k in range(3):
        np.testing.assert_allclose(res["marginals"][(k, k)]["ys"], od[:, k], rtol=tol, atol=1e-300)
    pi = 0
    for a in range(3):
        for b in range(a + 1, 3):
            np.testing.assert_allclose(res["marginals"][(a, b)]["ys"], op[:, :, pi], rtol=tol, atol=1e-300)
            np.testing.assert_array_equal(res["marginals"][(b, a)]["ys"], res["marginals"][(a, b)]["ys"].T)
            pi += 1
    # chunk boundaries never split a segment of G points
    try:
        ctx.set_chunk(128)
        r2 = bb.compute_marginals_int(prob, func=func, grid_size=gs, xs=lhc, normalize=False)
        for kk in res["marginals"]:
            np.testing.assert_array_equal(r2["marginals"][kk]["ys"], res["marginals"][kk]["ys"])
    finally:
        ctx.set_chunk(0)
    # normalisation (normalize_prob_vals!, ext/CairoExt.jl:410-430): trapezoid mass 1 on the plotting grid
    rn = bb.compute_marginals_int(prob, func=func, grid_size=gs, xs=lhc)
    step = (pr.ub - pr.lb) / (gs - 1)
    ys = rn["marginals"][(0, 0)]["ys"]
    assert (ys.sum() - 0.5 * (ys[0] + ys[-1])) * step[0] == pytest.approx(1.0, rel=1e-12)


def test_concurrent_calls_on_one_context(bb, ctx):
    """SURVEY.md section 8b, threading: BOSS may call the closures from several threads (parallel multistart, threaded candidate
    evaluation).  Calls on one context serialise on its mutex; results are bit-identical to the serial ones."""
    import threading
    pr, model, like, prior, *_ = problem(bb, "C2")
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    sets = [queries(pr, 3000 + 517 * i, seed=40 + i) for i in range(6)]
    serial = [bb.posterior_eval(post, like, prior, Xq, 7, acq=bb.MaxVar()) for Xq in sets]
    out, errs = [None] * len(sets), []

    def work(i):
        try:
            for _ in range(3):
                out[i] = bb.posterior_eval(post, like, prior, sets[i], 7, acq=bb.MaxVar())
                k = bb.Kernel(bb.SE, np.ones(pr.d))
                bb.calc_mmd(k, sets[i][:, :400], sets[i][:, 400:900], ctx)          # a different entry point in between
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(sets))]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    for r, s in zip(out, serial):
        for key in ("log_approx_post", "log_post_mean", "log_post_var"):
            assert np.array_equal(r[key], s[key], equal_nan=True)
        assert r["best_idx"] == s["best_idx"] and r["best_val"] == s["best_val"]
