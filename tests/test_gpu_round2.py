"""GPU: round-2 additions through the C ABI -- stream ordering of device inputs, multi-GPU from one process
(bosip_b200_init_multi and the `_multi` entry points), the MultiFittedParams variants of the argmax / evidence sweeps, the
INT8 engine's two-issuer self-check, and the guards around prior mean functions."""
import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu


def _c2(bb):
    pr = bb.synth.make_problem("C2")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    return pr, model, bb.NormalLikelihood(pr.z_obs, pr.std_obs), bb.ProductPrior(pr.prior_kind, pr.prior_params)


def test_device_inputs_are_ordered_after_the_torch_stream(bb, ctx):
    """the library's compute stream is non-blocking, so work torch has only ENQUEUED on its own stream -- a
    long kernel followed by the asynchronous fill of the query tensor, and the row-major -> column-major copy `_colmajor` itself
    issues -- must be ordered before the library reads the raw pointers.  Without bosip_b200_stream_wait this reads stale NaNs."""
    import torch
    pr, model, like, prior = _c2(bb)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    M = 3_000_000
    rng = np.random.default_rng(3)
    Xh = torch.as_tensor(rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, M))).pin_memory()     # ROW-major (d, M)
    ref = bb.posterior_eval(post, like, prior, Xh.numpy(), 1)["log_approx_post"]
    for trial in range(3):
        Xd = torch.full((pr.d, M), float("nan"), dtype=torch.float64, device="cuda")
        big = torch.randn(8192, 8192, device="cuda")
        torch.cuda.synchronize()
        for _ in range(12):                       # tens of ms of queued work on torch's current stream ...
            big = (big @ big) * 1e-4
        Xd.copy_(Xh, non_blocking=True)           # ... then the fill of the inputs, still only enqueued when the library is called
        got = bb.posterior_eval(post, like, prior, Xd, 1)["log_approx_post"]
        assert not torch.isnan(got).any(), f"trial {trial}: the library read the query tensor before torch had written it"
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
        # outputs handed in by the caller: memory torch's caching allocator may have recycled from a tensor whose kernels are
        # still queued (here: the matmul chain's temporaries)
        out = {"log_approx_post": torch.empty(M, dtype=torch.float64, device="cuda")}
        bb.posterior_eval(post, like, prior, Xd, 1, out=out)
        np.testing.assert_array_equal(out["log_approx_post"].cpu().numpy(), ref)
    post.free()


def test_multi_gpu_c_abi_matches_single_context(bb, ctx):
    """bosip_b200_init_multi + the `_multi` entry points (csrc/multi.cu): one process, one host thread per device.  Runs on every
    visible GPU (a one-GPU box exercises the same code with ndev = 1, plus a two-context split of the same device, which shards
    and combines exactly like two GPUs).  Per-point outputs and the argmax must be BIT-identical to a single context."""
    import torch
    pr, model, like, prior = _c2(bb)
    ndev = torch.cuda.device_count()
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    rng = np.random.default_rng(9)
    M = 100_003                                                     # not a multiple of anything
    Xq = rng.uniform(pr.lb[:, None] - 1.0, pr.ub[:, None] + 1.0, size=(pr.d, M))
    ref = bb.posterior_eval(post, like, prior, Xq, 7, acq=bb.LogMaxVar(), index_offset=17)
    src = bb.CandidateSource.philox(11, pr.lb, pr.ub, 250_001)
    ref_arg = bb.acq_argmax(post, like, prior, bb.MaxVar(), src)
    ref_pts = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(Xq))
    xs = prior.rand(20_000, np.random.default_rng(1))
    prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx); prob._post = post
    ref_ev = bb.evidence(prob, "mean", xs=xs)
    A = rng.standard_normal((3, 5000)); B = 0.3 + rng.standard_normal((3, 4000)); kern = bb.Kernel(bb.MATERN32, np.array([1.0, 2.0, 0.7]))
    ref_mmd = bb.calc_mmd(kern, A, B, ctx)
    G = 300_001; lw = rng.normal(0, 0.1, G); a = rng.normal(-3, 2, G); t = a + rng.normal(0, 0.5, G)
    ref_tv = bb.calc_tv(lw, a, t, ctx)
    layouts = [list(range(ndev))] if ndev > 1 else []
    layouts += [[0], [0, 0], [0, 0, 0]]                              # several contexts on one device: same sharding logic
    for devs in layouts:
        mctx = bb.MultiContext(devs)
        assert mctx.size == len(devs)
        mpost = bb.model_posterior_multi(model, pr.X, pr.Y, mctx)
        res = bb.posterior_eval_multi(mpost, like, prior, Xq, 7, acq=bb.LogMaxVar(), index_offset=17)
        for k in ("log_approx_post", "log_post_mean", "log_post_var"):
            np.testing.assert_array_equal(res[k], ref[k])
        assert (res["best_idx"], res["best_val"], res["n_clamped"]) == (ref["best_idx"], ref["best_val"], ref["n_clamped"])
        got = bb.acq_argmax_multi(mpost, like, prior, bb.MaxVar(), src)
        assert got[0] == ref_arg[0] and got[1] == ref_arg[1] and np.array_equal(got[2], ref_arg[2])
        got = bb.acq_argmax_multi(mpost, like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(Xq))
        assert got[0] == ref_pts[0] and got[1] == ref_pts[1] and np.array_equal(got[2], ref_pts[2])
        assert bb.evidence_multi(mpost, like, prior, xs, "mean") == pytest.approx(ref_ev, rel=1e-13)
        assert mctx.calc_mmd(kern, A, B) == pytest.approx(ref_mmd, rel=1e-12, abs=1e-15)
        assert mctx.calc_tv(lw, a, t) == pytest.approx(ref_tv, rel=1e-12)
        # host-supplied log-prior values are sharded with the points
        lp = prior.logpdf(Xq)
        r2 = bb.posterior_eval_multi(mpost, like, None, Xq, 2, logprior=lp)
        np.testing.assert_array_equal(r2["log_post_mean"], bb.posterior_eval(post, like, None, Xq, 2, logprior=lp)["log_post_mean"])
        # errors come back with the failing device's message
        with pytest.raises(bb.BosipB200Error):
            bb.posterior_eval_multi(mpost, bb.NormalLikelihood(pr.z_obs[:1], pr.std_obs[:1]), prior, Xq, 1)
        mpost.free(); mctx.close()
    post.free()


def test_bi_ensembles_through_the_high_level_paths(bb, ctx):
    """MultiFittedParams through B200CandidateAM / acq_argmax / evidence (bosip_b200_acq_argmax_bi,
    bosip_b200_evidence_sum_bi) against the oracle's log-mean-exp restatement (src/posterior/log_posterior.jl:86-95,153-162)."""
    pr, model, like, prior = _c2(bb)
    rng = np.random.default_rng(21)
    models, ofits = [], []
    for _ in range(3):
        lam = pr.lam * rng.uniform(0.7, 1.4, pr.lam.shape); alpha = pr.alpha * rng.uniform(0.8, 1.2, pr.p); sigma = pr.sigma * rng.uniform(0.5, 2.0, pr.p)
        models.append(bb.GaussianProcess(pr.kernel_id, lam, alpha, sigma))
        ofits.append(orc.gp_fit(pr.kernel_id, pr.X, pr.Y, lam, alpha, sigma))
    prob = bb.BosipProblem(pr.X, pr.Y, models, like, prior, (pr.lb, pr.ub), ctx)
    olike = orc.NormalLikelihood(pr.z_obs, pr.std_obs); oprior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    src = bb.CandidateSource.philox(4, pr.lb, pr.ub, 20_000)
    cand = bb.synth.philox_uniform_candidates(4, 0, 20_000, pr.lb, pr.ub)
    ores = orc.posterior_eval_bi(ofits, olike, oprior, cand)
    oi, ov = orc.maxvar_argmax(ores["log_post_var"], True)
    x, val = bb.B200CandidateAM(src).maximize_acquisition(prob, bb.LogMaxVar())
    assert np.array_equal(x, cand[:, oi]) and val == pytest.approx(ov, rel=1e-9)
    pts = bb.acq_argmax(prob.model_posterior(), like, prior, bb.LogMaxVar(), bb.CandidateSource.from_points(cand))
    assert pts[0] == oi
    xs = prior.rand(8000, np.random.default_rng(2))
    o2 = orc.posterior_eval_bi(ofits, olike, oprior, xs)
    for which, key in (("approx", "log_approx_post"), ("mean", "log_post_mean")):
        assert bb.evidence(prob, which, xs=xs) == pytest.approx(orc.evidence(o2[key], orc.log_prior(oprior, xs)), rel=1e-10)
    f = bb.approx_posterior(prob, normalize=True, xs=xs)            # the normalised closures need the ensemble evidence
    assert np.all(np.isfinite(f(xs[:, :10])))


def test_prior_mean_guards(bb, ctx):
    """a GP prior mean is a host callable -- device inputs must fail loudly instead of silently dropping it,
    and an ensemble whose samples carry different mean functions is rejected."""
    import torch
    pr, model, like, prior = _c2(bb)
    mean = lambda X: np.vstack([0.3 * np.asarray(X)[0], -0.2 * np.asarray(X)[1]])      # noqa: E731
    m1 = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma, mean=mean)
    post = bb.model_posterior(m1, pr.X, pr.Y, ctx=ctx)
    Xq = np.random.default_rng(0).uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, 500))
    host = bb.posterior_eval(post, like, prior, Xq, 3)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma, mean_at_X=mean(pr.X))
    ores = orc.posterior_eval(ofit, orc.NormalLikelihood(pr.z_obs, pr.std_obs), orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params]),
                              Xq, mean_at_Xq=mean(Xq))
    np.testing.assert_allclose(host["log_post_mean"], ores["log_post_mean"], rtol=1e-10)
    Xd = torch.as_tensor(Xq.T.copy(), device="cuda").t()
    with pytest.raises(NotImplementedError):
        bb.posterior_eval(post, like, prior, Xd, 3)
    with pytest.raises(NotImplementedError):
        post.mean(Xd)
    m2 = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma, mean=lambda X: 0.0 * mean(X))
    post2 = bb.model_posterior(m2, pr.X, pr.Y, ctx=ctx)
    with pytest.raises(NotImplementedError):
        bb.posterior_eval([post, post2], like, prior, Xq, 3)
    both = bb.posterior_eval([post, post], like, prior, Xq, 3)       # same mean function on every sample: fine
    np.testing.assert_allclose(both["log_post_mean"], host["log_post_mean"], rtol=1e-12)


def test_int8_two_issuer_self_check_ran(bb):
    """the two-issuer INT8 kernel is used only after it reproduced the single-issuer kernel bit for bit on the
    context's first real launch (vargemm_i8.cu, launch_i8)."""
    pr, model, like, prior = _c2(bb)
    c = bb.Context(0)
    assert c.i8_issuers() == 0
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=c)
    Xq = np.random.default_rng(1).uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, 3000))
    c.set_variance_engine("int8")
    first = bb.posterior_eval(post, like, prior, Xq, 4)["log_post_var"]          # this call carries the self-check
    assert c.i8_issuers() in (1, 2)
    again = bb.posterior_eval(post, like, prior, Xq, 4)["log_post_var"]
    np.testing.assert_array_equal(first, again)                                    # the checked launch already returned final values
    post.free(); c.close()
