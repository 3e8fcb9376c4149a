"""GPU: the AMIS inner loop, down-sampling, confidence-set and OptMMD pieces through the C ABI
vs the oracle restatement of src/samplers/amis/amis_method.jl, proposal_distributions/normal.jl, src/sampling.jl,
src/utils/confidence_set.jl and src/term_conds/ae_confidence.jl."""
import math

import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _problem(bb, ctx):
    pr = bb.synth.make_problem("C2")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    return pr, bb.BosipProblem(pr.X, pr.Y, model, like, prior, bounds=(pr.lb, pr.ub), ctx=ctx)


def test_mvnormal_draws_match_host_mirror_and_moments(bb, ctx):
    rng = np.random.default_rng(0)
    for d in (1, 2, 5, 8):
        A = rng.standard_normal((d, d)); Sigma = A @ A.T + 0.5 * np.eye(d); mu = rng.standard_normal(d)
        q = bb.NormalProposal(mu, Sigma)
        x = q.rand(4000, seed=77, offset=123, ctx=ctx)
        # host mirror: same Philox uniforms, Box-Muller, x = mu + L z
        d2 = d + (d & 1)
        u = bb.synth.philox_uniform_candidates(77, 123, 4000, np.zeros(d2), np.ones(d2))
        z = np.empty((d2, 4000))
        for c in range(0, d2, 2):
            r = np.sqrt(-2.0 * np.log(1.0 - u[c]))
            z[c] = r * np.cos(2 * np.pi * u[c + 1]); z[c + 1] = r * np.sin(2 * np.pi * u[c + 1])
        ref = mu[:, None] + np.linalg.cholesky(Sigma) @ z[:d]
        assert np.max(np.abs(x - ref)) < 1e-11 * (1 + np.max(np.abs(ref)))
        # draws depend only on (seed, offset + i): a shifted window reproduces the overlap; device buffers agree with host ones
        x2 = q.rand(1000, seed=77, offset=1123, ctx=ctx)
        assert np.array_equal(x2, x[:, 1000:2000])
        xd = q.rand(4000, seed=77, offset=123, ctx=ctx, device=True)
        assert np.array_equal(xd.cpu().numpy(), x)
    big = bb.NormalProposal(mu, Sigma).rand(400000, seed=5, ctx=ctx)
    assert np.max(np.abs(big.mean(axis=1) - mu)) < 0.02 and np.max(np.abs(np.cov(big) - Sigma)) < 0.05 * np.max(np.abs(Sigma))


def test_mixture_pdf_sums_weights_and_moments_vs_oracle(bb, ctx):
    import torch
    rng = np.random.default_rng(1)
    d, M = 5, 7001
    qs = []
    for _ in range(9):
        A = rng.standard_normal((d, d)); qs.append(bb.NormalProposal(rng.standard_normal(d), A @ A.T + 0.3 * np.eye(d)))
    X = rng.standard_normal((d, M)) * 1.5
    ref = sum(np.exp(orc.mvnormal_logpdf(q.mu, q.Sigma, X)) for q in qs)
    delta = bb.mixture_pdf_sum(qs, X, ctx=ctx)
    np.testing.assert_allclose(delta, ref, rtol=1e-12)
    delta2 = bb.mixture_pdf_sum(qs[:3], X, delta=delta.copy(), ctx=ctx)                  # accumulate
    np.testing.assert_allclose(delta2, ref + sum(np.exp(orc.mvnormal_logpdf(q.mu, q.Sigma, X)) for q in qs[:3]), rtol=1e-12)
    Xd = torch.as_tensor(np.ascontiguousarray(X.T), device="cuda").t()
    dd = bb.mixture_pdf_sum(qs, Xd, ctx=ctx)
    assert dd.is_cuda and np.array_equal(dd.cpu().numpy(), delta)                         # host and device buffers: same bits
    many = [qs[i % 9] for i in range(150)]                                                # more proposals than one launch holds
    np.testing.assert_allclose(bb.mixture_pdf_sum(many, X, ctx=ctx), sum(np.exp(orc.mvnormal_logpdf(q.mu, q.Sigma, X)) for q in many), rtol=1e-12)
    # log-weights and the AnalyticalFitter sums
    log_p = -0.5 * np.sum(X * X, axis=0) + rng.normal(0, 0.1, M)
    lw = bb.amis_log_weights(log_p, delta, 9.0, ctx=ctx)
    np.testing.assert_allclose(lw, log_p - np.log(ref / 9.0), rtol=1e-12, atol=1e-12)
    lw[::50] = -np.inf                                                                    # zero weights are legal
    mean, cov, V1, V2, ws = bb.weighted_moments(X, lw, want_ws=True, ctx=ctx)
    ows = orc.get_weights(lw)
    np.testing.assert_allclose(ws, ows, rtol=1e-12, atol=0)
    omu, oS = orc.estimate_parameters(X, ows)
    np.testing.assert_allclose(mean, omu, rtol=1e-11, atol=1e-13); np.testing.assert_allclose(cov, oS, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(bb.exp_weights(lw, ctx=ctx), ows, rtol=1e-12)
    # error behaviour of get_weights / exp_weights (amis_method.jl:112-114,128)
    for bad in (np.nan, np.inf):
        l2 = lw.copy(); l2[3] = bad
        with pytest.raises(RuntimeError, match="Numerical issues"):
            bb.weighted_moments(X, l2, ctx=ctx)
    with pytest.raises(RuntimeError, match="Numerical issues"):
        bb.weighted_moments(X, np.full(M, -np.inf), ctx=ctx)


@pytest.mark.parametrize("device", [True, False])
def test_amis_replay_matches_oracle(bb, ctx, device):
    """Full AMIS run on the device; the oracle replays the reference's update rules on the same draws."""
    pr, problem = _problem(bb, ctx)
    log_pi = bb.log_posterior_mean(problem)
    q0 = bb.NormalProposal(0.5 * (pr.lb + pr.ub), np.diag((pr.ub - pr.lb) / 5.0))
    amis = bb.AMIS(T=6, N=400)
    xs, ws, tr = amis(log_pi, q0, bb.AnalyticalFitter(), seed=11, ctx=ctx, device=device, return_trace=True)
    host = (lambda v: v.cpu().numpy()) if device else (lambda v: np.asarray(v))
    xs_all, log_P = host(tr["xs"]), host(tr["log_P"])
    # log_pi of the draws equals the oracle's closure on the same points
    post = problem.model_posterior()
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    ores = orc.posterior_eval(ofit, orc.NormalLikelihood(pr.z_obs, pr.std_obs),
                              orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params]), xs_all)
    fin = np.isfinite(ores["log_post_mean"])
    assert np.array_equal(np.isneginf(log_P), np.isneginf(ores["log_post_mean"]))
    np.testing.assert_allclose(log_P[fin], ores["log_post_mean"][fin], rtol=TOL)
    qs, Delta, log_O, ows = orc.amis_replay(xs_all, log_P, 400, (q0.mu, q0.Sigma))
    for t, (mu, Sig) in enumerate(qs):
        np.testing.assert_allclose(tr["qs"][t].mu, mu, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(tr["qs"][t].Sigma, Sig, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(host(tr["Delta"]), Delta, rtol=1e-9)
    lo = host(tr["log_Omega"]); f2 = np.isfinite(log_O)
    assert np.array_equal(np.isfinite(lo), f2)
    np.testing.assert_allclose(lo[f2], log_O[f2], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(host(ws), ows, rtol=1e-8, atol=1e-300)
    assert host(xs).shape == (pr.d, 6 * 400) and abs(float(host(ws).sum()) - 1.0) < 1e-12
    # the proposal has moved towards the posterior mass: effective sample size is a sane fraction of the draws
    ess = 1.0 / np.sum(host(ws) ** 2)
    assert ess > 50


def test_resample_and_sample_posterior_pure(bb, ctx):
    rng = np.random.default_rng(2)
    d, M, count = 3, 50000, 200000
    X = rng.standard_normal((d, M)); w = rng.gamma(0.3, 1.0, M); w[::7] = 0.0
    out, idx = bb.resample(X, w, count, seed=9, ctx=ctx, return_index=True)
    assert out.shape == (d, count) and np.array_equal(out, X[:, idx])
    assert np.all(w[idx] > 0)                                         # zero-weight columns are never drawn
    # host mirror: same uniforms, bisection on the running sum (block-wise partial sums may move a boundary by an ulp)
    u = bb.synth.philox_uniform_candidates(9, 0, count, np.zeros(2), np.ones(2))[0]
    cdf = np.cumsum(w)
    ref = np.searchsorted(cdf, u * cdf[-1], side="right")
    assert np.mean(ref == idx) > 0.9999
    # frequencies follow the weights: chi-square on 20 weight-quantile bins
    order = np.argsort(w); bins = np.array_split(order, 20)
    exp = np.array([w[b].sum() for b in bins]) / w.sum() * count
    obs = np.array([np.isin(idx, b).sum() for b in bins])
    keep = exp > 5
    assert np.sum((obs[keep] - exp[keep]) ** 2 / exp[keep]) < 60.0
    # sample_posterior_pure: AMIS + down-sampling end to end, samples stay on the device and feed calc_mmd directly
    pr, problem = _problem(bb, ctx)
    sampler = bb.AMISSampler(iters=5)
    xs = bb.sample_posterior_pure(sampler, bb.log_posterior_mean(problem), (pr.lb, pr.ub), 300, supersample_ratio=4, seed=3, ctx=ctx)
    assert xs.is_cuda and tuple(xs.shape) == (pr.d, 300)
    xh = xs.cpu().numpy()
    assert np.all(xh >= pr.lb[:, None]) and np.all(xh <= pr.ub[:, None])    # posterior mass lies inside the prior support
    m = bb.calc_mmd(bb.Kernel(bb.SE, np.ones(pr.d)), xs, xs, ctx)
    assert abs(m) < 1e-12


def test_confidence_set_device_kernels_edge_cases(bb, ctx):
    """csrc/confset.cu against the oracle's loop-for-loop restatement of src/utils/confidence_set.jl:19-81 (+ StatsBase's weighted
    quantile): tiny and ragged sizes around the 2048-element sort tile, ties, zero weights, NaN, device-resident inputs."""
    import torch
    rng = np.random.default_rng(12)
    for n in (1, 2, 3, 17, 400, 2047, 2048, 2049, 5000, 70001):
        v = rng.normal(size=n); w = rng.uniform(0.0, 2.0, n); w[rng.random(n) < 0.2] = 0.0
        if not np.any(w):
            w[0] = 1.0
        for p in (0.0, 0.03, 0.5, 0.8):
            assert bb.weighted_quantile(v, w, p, ctx) == pytest.approx(orc.weighted_quantile(v, w, p), rel=1e-12, abs=1e-15)
        # p = 1: h = (W - w_1) + w_1 lands within rounding of the total weight, so whether the last cumulative sum exceeds it (and
        # the last two values are interpolated with a fraction 1 - O(n eps)) depends on the summation order, in the reference too:
        # the result is the largest kept value up to O(n eps W / w_last) of the last gap
        top = np.max(v[w > 0])
        assert bb.weighted_quantile(v, w, 1.0, ctx) == pytest.approx(top, abs=1e-9 * (np.ptp(v) + 1e-300) + 1e-15)
        assert orc.weighted_quantile(v, w, 1.0) == pytest.approx(top, abs=1e-9 * (np.ptp(v) + 1e-300) + 1e-15)
        v[n // 2] = v[0]                                              # ties
        assert bb.weighted_quantile(v, w, 0.37, ctx) == pytest.approx(orc.weighted_quantile(v, w, 0.37), rel=1e-12, abs=1e-15)
        vd, wd = torch.as_tensor(v, device="cuda"), torch.as_tensor(w, device="cuda")
        assert bb.weighted_quantile(vd, wd, 0.37, ctx) == bb.weighted_quantile(v, w, 0.37, ctx)      # host and device buffers: same bits
        assert bb.weighted_quantile(v, None, 0.25, ctx) == pytest.approx(orc.weighted_quantile(v, np.ones(n), 0.25), rel=1e-12, abs=1e-15)
    xs = rng.normal(size=(2, 500)); pdf = lambda X: np.exp(-0.5 * (np.asarray(X) ** 2).sum(0)); ws = rng.uniform(0.5, 1.5, 500)   # noqa: E731
    vals = pdf(xs)
    for q in (0.5, 0.8, 0.95):
        c = bb.find_cutoff(pdf, xs, ws, q, ctx)
        assert c == pytest.approx(orc.find_cutoff(vals, ws, q), rel=1e-12)
        assert bb.find_cutoff(pdf, xs, q, ctx=ctx) == pytest.approx(orc.find_cutoff(vals, np.ones(500), q), rel=1e-12)
        assert bb.approx_cutoff_area(pdf, xs, ws, c, ctx) == pytest.approx(orc.approx_cutoff_area(vals, ws, c), rel=1e-12)
        assert bb.approx_cutoff_area(pdf, xs, ws, c, ctx) == pytest.approx(q, abs=0.02)        # the cut-off encloses ~q of the weight
        assert bb.approx_cutoff_area(pdf, xs, c, ctx=ctx) == pytest.approx(orc.approx_cutoff_area(vals, np.ones(500), c), rel=1e-12)
    in_a = vals >= np.quantile(vals, 0.3); in_b = pdf(xs - 0.4) >= np.quantile(pdf(xs - 0.4), 0.3)
    lp = -0.5 * (xs ** 2).sum(0) / 9.0
    assert bb.set_iou(in_a, in_b, lp, ctx) == pytest.approx(orc.set_iou(in_a, in_b, np.exp(lp)), rel=1e-12)
    assert bb.set_iou(in_a, in_a, lp, ctx) == pytest.approx(1.0, rel=1e-14)
    assert np.isnan(bb.weighted_quantile(np.array([1.0, np.nan]), np.ones(2), 0.5, ctx))
    # the fused AEConfidence core on log-values: host buffers == device buffers == the oracle's three steps
    n = 30011
    la = rng.normal(-3.0, 2.0, n); lb = la + rng.normal(0.0, 0.3, n); lpr = rng.normal(-2.0, 0.5, n)
    iou, ca, cb = bb.confidence_iou(la, lb, lpr, 0.9, ctx)
    va, vb = np.exp(la), np.exp(lb)
    oca = orc.find_cutoff(va, np.exp(np.log(va) - lpr), 0.9); ocb = orc.find_cutoff(vb, np.exp(np.log(vb) - lpr), 0.9)
    assert ca == pytest.approx(oca, rel=1e-12) and cb == pytest.approx(ocb, rel=1e-12)
    assert iou == pytest.approx(orc.set_iou(va > oca, vb > ocb, np.exp(lpr)), rel=1e-10)
    dev = bb.confidence_iou(torch.as_tensor(la, device="cuda"), torch.as_tensor(lb, device="cuda"), lpr, 0.9, ctx)
    assert dev == (iou, ca, cb)


def test_confidence_sets_and_ae_confidence_vs_oracle(bb, ctx):
    rng = np.random.default_rng(4)
    vals = rng.gamma(2.0, 1.0, 5000); ws = rng.gamma(1.0, 1.0, 5000); ws[::11] = 0.0
    for p in (0.0, 0.05, 0.5, 0.95, 1.0):
        assert bb.weighted_quantile(vals, ws, p, ctx) == pytest.approx(orc.weighted_quantile(vals, ws, p), rel=1e-12)
    assert bb.weighted_quantile([1, 2, 3, 4], np.ones(4), 0.5, ctx) == 2.5     # StatsBase: h = p (W - w1) + w1 on the cumulative weights
    xs = rng.uniform(-1, 1, (2, 5000))
    f = lambda X: np.exp(-np.sum(np.asarray(X) ** 2, axis=0))             # noqa: E731
    c = bb.find_cutoff(f, xs, ws, 0.8, ctx)
    assert c == pytest.approx(orc.find_cutoff(f(xs), ws, 0.8), rel=1e-12)
    assert bb.find_cutoff(f, xs, 0.8, ctx=ctx) == pytest.approx(orc.find_cutoff(f(xs), np.ones(5000), 0.8), rel=1e-12)
    assert bb.approx_cutoff_area(f, xs, ws, c, ctx) == pytest.approx(orc.approx_cutoff_area(f(xs), ws, c), rel=1e-12)
    inA, inB = f(xs) > c, f(xs + 0.1) > c
    lp = rng.normal(-1.0, 0.3, 5000)
    assert bb.set_iou(inA, inB, lp, ctx) == pytest.approx(orc.set_iou(inA, inB, np.exp(lp)), rel=1e-12)
    # AEConfidence.calculate: both posteriors evaluated by the fused device sweep, then the reference's quantile / IoU steps
    pr, problem = _problem(bb, ctx)
    xs = problem.x_prior.rand(6000, np.random.default_rng(8))
    cond = bb.AEConfidence(xs=xs, q=0.9, r=0.95)
    iou = cond.calculate(problem)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    ores = orc.posterior_eval(ofit, orc.NormalLikelihood(pr.z_obs, pr.std_obs), oprior, xs)
    va, ve = np.exp(ores["log_approx_post"]), np.exp(ores["log_post_mean"])
    olp = orc.log_prior(oprior, xs)
    np.testing.assert_allclose(problem.x_prior.logpdf(xs), olp, rtol=1e-13)
    ca = orc.find_cutoff(va, np.exp(np.log(va) - olp), 0.9); ce = orc.find_cutoff(ve, np.exp(np.log(ve) - olp), 0.9)
    ref = orc.set_iou(va > ca, ve > ce, np.exp(olp))
    assert iou == pytest.approx(ref, rel=1e-6) and 0.0 < iou <= 1.0
    assert cond(problem) == (iou < 0.95)


def test_opt_mmd_metric(bb, ctx):
    rng = np.random.default_rng(6)
    A = rng.standard_normal((2, 1500)); B = 0.4 + 1.2 * rng.standard_normal((2, 1500))
    metric = bb.OptMMDMetric(kernel=bb.SE, bounds=(np.array([-5.0, -5.0]), np.array([5.0, 5.0])), options={"maxiter": 60, "xatol": 1e-3, "fatol": 1e-9})
    val = metric.calculate(A, B, ctx)
    lam0 = np.full(2, 10.0 / 3.0)
    start = bb.calc_mmd(bb.Kernel(bb.SE, lam0), A, B, ctx)
    assert val >= start - 1e-15                                               # the optimiser maximises from x0 = (ub - lb) / 3
    assert val == pytest.approx(orc.calc_mmd(orc.SE, metric.lengthscales_, A, B)[0], rel=1e-9)
    assert np.all(metric.lengthscales_ > 0) and np.all(metric.lengthscales_ <= 10.0)
