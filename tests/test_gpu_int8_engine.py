"""GPU: the INT8 (tcgen05, error-free byte slicing) variance engine through the C ABI vs the oracle and vs the FP64 (DMMA)
engine: same parity gates as tests/test_gpu_parity.py, ragged shapes around the 64/128 tile edges, chunk invariance,
queries on top of training points (worst cancellation), BI ensembles, option errors."""
import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture()
def engine(ctx):
    """Yields a setter; always restores the default (auto) engine."""
    yield ctx.set_variance_engine
    ctx.set_variance_engine("auto")


def _var_err(var, ovar, kxx):
    return float(np.max(np.abs(var - ovar) / (np.abs(ovar) + 1e-3 * kxx[:, None])))


def _problem(bb, name, M, seed=5):
    pr = bb.synth.make_problem(name)
    rng = np.random.default_rng(seed)
    Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, M))
    n = min(48, pr.N)
    Xq[:, :n] = pr.X[:, :n]                 # exactly on training points: sigma^2 = alpha^2 - q cancels almost completely
    Xq[:, n:2 * n] = pr.X[:, :n] + 1e-7
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    return pr, model, ofit, Xq


@pytest.mark.parametrize("name,M", [("C2", 20000), ("C3", 20000)])
@pytest.mark.parametrize("slices", [0, 7, 8])      # 0 = the default: 7 slices of alpha^2 L^-1 against 6 of K* (27 slice pairs)
def test_int8_engine_vs_oracle_and_fp64(bb, ctx, engine, name, M, slices):
    pr, model, ofit, Xq = _problem(bb, name, M)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    omu, ovar = orc.gp_predict(ofit, Xq)
    engine("fp64")
    mu0, var0 = post.mean_and_var(Xq)
    engine("int8", slices)
    mu1, var1 = post.mean_and_var(Xq)
    kxx = pr.alpha ** 2
    assert np.array_equal(mu0, mu1)                       # the mean path does not depend on the engine
    assert _var_err(var1, ovar, kxx) < TOL
    assert _var_err(var1, var0, kxx) < TOL
    if slices == 0:
        assert ctx.variance_engine() == {"engine": "int8", "slices": 7, "kstar_slices": 6, "slice_pairs": 27}
        assert _var_err(var1, ovar, kxx) < 10 * _var_err(var0, ovar, kxx) + 5e-12  # a few times the Float64 engine's error, >= 10 x below the gate
    else:
        assert _var_err(var1, ovar, kxx) < 4 * _var_err(var0, ovar, kxx) + 1e-12   # as accurate as plain Float64
    # fused closures on the INT8 engine
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    res = bb.posterior_eval(post, like, prior, Xq, 7)
    ores = orc.posterior_eval(ofit, orc.NormalLikelihood(pr.z_obs, pr.std_obs),
                              orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params]), Xq)
    fin = np.isfinite(ores["log_post_var"]) & np.isfinite(res["log_post_var"])
    assert np.sum(np.isfinite(ores["log_post_var"]) != np.isfinite(res["log_post_var"])) <= 1e-3 * M + 1
    assert np.all(np.abs(res["log_post_var"][fin] - ores["log_post_var"][fin]) <= TOL * np.abs(ores["log_post_var"][fin]) + 1e-7)
    np.testing.assert_allclose(res["log_post_mean"][np.isfinite(ores["log_post_mean"])],
                               ores["log_post_mean"][np.isfinite(ores["log_post_mean"])], rtol=TOL)


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
@pytest.mark.parametrize("d,p,N,M", [(1, 1, 1, 1), (3, 2, 63, 127), (3, 2, 64, 128), (4, 3, 65, 129), (2, 1, 127, 300), (2, 2, 128, 256),
                                       (6, 1, 129, 1000), (5, 2, 257, 500), (10, 1, 400, 777)])
@pytest.mark.parametrize("slices", [0, 7])
def test_int8_ragged_shapes(bb, ctx, engine, kernel_id, d, p, N, M, slices):
    rng = np.random.default_rng(100 * kernel_id + N)
    X = rng.uniform(-5, 5, (d, N)); Y = rng.standard_normal((p, N))
    lam = rng.uniform(1, 4, (d, p)); alpha = rng.uniform(0.5, 2, p); sigma = 0.05 * alpha
    Xq = rng.uniform(-5, 5, (d, M))
    model = bb.GaussianProcess(kernel_id, lam, alpha, sigma)
    post = bb.model_posterior(model, X, Y, ctx=ctx)
    ofit = orc.gp_fit(kernel_id, X, Y, lam, alpha, sigma)
    omu, ovar = orc.gp_predict(ofit, Xq)
    engine("int8", slices)
    mu, var = post.mean_and_var(Xq)
    assert var.shape == (p, M)
    assert _var_err(var, ovar, alpha ** 2) < TOL
    scale = np.max(np.abs(omu))
    assert np.max(np.abs(mu - omu) / (np.abs(omu) + 1e-3 * scale)) < 1e-9


def test_int8_chunking_is_invisible(bb, ctx, engine):
    pr, model, ofit, Xq = _problem(bb, "C2", 3001)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    engine("int8", 0)
    mu0, var0 = post.mean_and_var(Xq)
    try:
        for chunk in (128, 640, 1000):
            ctx.set_chunk(chunk)
            mu, var = post.mean_and_var(Xq)
            assert np.array_equal(mu, mu0) and np.array_equal(var, var0)
    finally:
        ctx.set_chunk(0)
    import torch
    Xd = torch.from_numpy(np.ascontiguousarray(Xq.T)).cuda().t()
    mu_d, var_d = post.mean_and_var(Xd)
    assert np.array_equal(mu_d.cpu().numpy(), mu0) and np.array_equal(var_d.cpu().numpy(), var0)


def test_auto_engine_and_option_errors(bb, ctx, engine):
    pr, model, ofit, Xq = _problem(bb, "C2", 2000)      # N = 200 >= 128: auto picks the INT8 engine
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    engine("auto")
    _, va = post.mean_and_var(Xq)
    engine("int8", 0)
    _, vi = post.mean_and_var(Xq)
    assert np.array_equal(va, vi)
    pr1, model1, ofit1, Xq1 = _problem(bb, "C1", 500)   # N = 20 < 128: auto stays on the FP64 engine
    post1 = bb.model_posterior(model1, pr1.X, pr1.Y, ctx=ctx)
    engine("auto")
    _, va = post1.mean_and_var(Xq1)
    engine("fp64")
    _, vf = post1.mean_and_var(Xq1)
    assert np.array_equal(va, vf)
    with pytest.raises(bb.BosipB200Error):
        ctx.set_variance_engine("int8", 5)
    with pytest.raises(bb.BosipB200Error):
        ctx.set_variance_engine("int8", 9)
    assert ctx.int8_peak() > 1000.0                      # TOPS; the tcgen05 path is alive


def test_int8_maximum_training_set_size(bb, ctx, engine):
    """The INT8 engine's exact integer accumulation holds for N < 8192 (int32 group sums, the int64 assembly): the largest supported
    size against the oracle (default 7 x 6 slicing and the symmetric 7 x 7 one), and the limit itself -- N = 8192 is refused on an
    explicit INT8 request and falls back to the FP64 engine under `auto`."""
    d, p, N, M = 6, 1, 8191, 300
    rng = np.random.default_rng(8191)
    X = rng.uniform(-5, 5, (d, N)); Y = np.sin(X).sum(axis=0)[None, :] + 0.05 * rng.standard_normal((p, N))
    lam = rng.uniform(1.5, 4, (d, p)); alpha = np.array([1.3]); sigma = 0.05 * alpha
    Xq = rng.uniform(-5, 5, (d, M)); Xq[:, :16] = X[:, :16]
    model = bb.GaussianProcess(2, lam, alpha, sigma)
    post = bb.model_posterior(model, X, Y, ctx=ctx)
    ofit = orc.gp_fit(2, X, Y, lam, alpha, sigma)
    omu, ovar = orc.gp_predict(ofit, Xq)
    kxx = alpha ** 2
    engine("fp64")
    _, var64 = post.mean_and_var(Xq)
    for slices in (0, 7):
        engine("int8", slices)
        mu, var = post.mean_and_var(Xq)
        assert _var_err(var, ovar, kxx) < TOL and _var_err(var, var64, kxx) < TOL
        assert _var_err(var, ovar, kxx) < 10 * _var_err(var64, ovar, kxx) + 5e-12
        assert np.max(np.abs(mu - omu)) <= 1e-9 * np.max(np.abs(ofit.a)) * N      # a cancelling sum of N terms, see the test above
    post.free()
    # the limit
    N2 = 8192
    X2 = rng.uniform(-5, 5, (2, N2)); Y2 = np.sin(X2).sum(axis=0)[None, :]
    model2 = bb.GaussianProcess(2, np.full((2, 1), 2.0), np.array([1.0]), np.array([0.05]))
    post2 = bb.model_posterior(model2, X2, Y2, ctx=ctx)
    Xq2 = rng.uniform(-5, 5, (2, 130))
    engine("int8", 0)
    with pytest.raises(bb.BosipB200Error):
        post2.mean_and_var(Xq2)
    engine("auto")
    _, va = post2.mean_and_var(Xq2)
    engine("fp64")
    _, vf = post2.mean_and_var(Xq2)
    assert np.array_equal(va, vf)
    post2.free()


def test_int8_bi_ensemble_matches_fp64(bb, ctx, engine):
    pr = bb.synth.make_problem("C2")
    rng = np.random.default_rng(3)
    models = [bb.GaussianProcess(pr.kernel_id, pr.lam * f, pr.alpha, pr.sigma) for f in (0.8, 1.0, 1.3)]
    posts = [bb.model_posterior(m, pr.X, pr.Y, ctx=ctx) for m in models]
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, 4000))
    engine("fp64")
    r0 = bb.posterior_eval(posts, like, prior, Xq, 7)
    engine("int8", 7)
    r1 = bb.posterior_eval(posts, like, prior, Xq, 7)
    np.testing.assert_array_equal(r0["log_approx_post"], r1["log_approx_post"])   # plug-in closure: no variance involved
    np.testing.assert_allclose(r0["log_post_mean"], r1["log_post_mean"], rtol=TOL)
    a, b = r0["log_post_var"], r1["log_post_var"]
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.sum(np.isfinite(a) != np.isfinite(b)) <= 4
    assert np.all(np.abs(a[fin] - b[fin]) <= TOL * np.abs(a[fin]) + 1e-7)


def test_two_mma_issuers_equal_one_bit_for_bit(bb, ctx, engine, tmp_path):
    """The default kernel deals the slice rows of a k-block to two MMA-issuing threads that accumulate into the same TMEM columns;
    integer accumulation makes the result independent of how the two streams interleave.  BOSIP_B200_I8_ISSUERS=1 (read at first use,
    hence the subprocess) selects the single-issuer kernel: the variances must agree to the last bit."""
    import os
    import subprocess
    import sys
    pr, model, ofit, Xq = _problem(bb, "C3", 6000)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    engine("int8", 7)
    _, var2 = post.mean_and_var(Xq)
    out = tmp_path / "var1.npy"
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {repr(os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))})\n"
        "sys.path.insert(0, sys.path[0] + '/tests')\n"
        "import bosip_b200 as bb\n"
        "from test_gpu_int8_engine import _problem\n"
        "ctx = bb.Context(0); ctx.set_variance_engine('int8', 7)\n"
        "pr, model, ofit, Xq = _problem(bb, 'C3', 6000)\n"
        "post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)\n"
        f"np.save({repr(str(out))}, post.mean_and_var(Xq)[1])\n")
    env = dict(os.environ, BOSIP_B200_I8_ISSUERS="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
    assert np.array_equal(np.load(out), var2)


@pytest.mark.parametrize("slices", [6, 7, 8])
def test_mma_schedules_are_bit_identical(bb, ctx, engine, tmp_path, slices):
    """The three MMA schedules of the INT8 kernel (BOSIP_B200_I8_MODE: 0 plain, 1 phased, 2 / 3 split last / first and last k-block; vargemm_i8.cu) only
    reorder exact integer accumulations: every slice count must give the same bits, on a shape with a ragged last tile and tiles of
    one to several k-blocks.  The mode is read at first use, hence the subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    outs = []
    for mode in (0, 1, 2, 3):
        out = tmp_path / f"var_m{mode}.npy"
        code = (
            "import sys, numpy as np\n"
            f"sys.path.insert(0, {repr(root)})\n"
            "import bosip_b200 as bb\n"
            f"ctx = bb.Context(0); ctx.set_variance_engine('int8', {slices})\n"
            "pr = bb.synth.make_problem('X', d=4, p=2, N=715, seed=77)\n"
            "model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)\n"
            "Xq = np.random.default_rng(5).uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, 3001))\n"
            "post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)\n"
            f"np.save({repr(str(out))}, post.mean_and_var(Xq)[1]); assert ctx.i8_issuers() == 2\n")
        subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, BOSIP_B200_I8_MODE=str(mode)), timeout=300)
        outs.append(np.load(out))
    assert all(np.array_equal(outs[0], o) for o in outs[1:])


def test_paired_kernel_is_bit_identical(bb, ctx, engine, tmp_path):
    """The paired INT8 kernel (tcgen05 cta_group::2: two CTAs of a cluster share one M = 256 MMA stream, per-rank operand images, relay
    and multicast barriers; opt-in, BOSIP_B200_I8_PAIR=1) must give the same bits as the unpaired kernel: odd and even numbers of query
    tiles, a ragged last n-tile, tiles of one to several k-blocks.  The switch is read at first use, hence the subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    outs = {}
    for pair in (0, 1):
        for M in (3001, 1100):          # 24 and 9 query tiles
            out = tmp_path / f"var_p{pair}_{M}.npy"
            code = (
                "import sys, numpy as np\n"
                f"sys.path.insert(0, {repr(root)})\n"
                "import bosip_b200 as bb\n"
                "ctx = bb.Context(0); ctx.set_variance_engine('int8', 7)\n"
                "pr = bb.synth.make_problem('X', d=4, p=2, N=715, seed=77)\n"
                "model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)\n"
                f"Xq = np.random.default_rng(5).uniform(pr.lb[:, None], pr.ub[:, None], size=(pr.d, {M}))\n"
                "post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)\n"
                "v1 = post.mean_and_var(Xq)[1]; v2 = post.mean_and_var(Xq)[1]\n"      # first call: self-check launch, second: the kernel alone
                f"assert np.array_equal(v1, v2); np.save({repr(str(out))}, v2); assert ctx.i8_pair() == {pair}\n")
            subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, BOSIP_B200_I8_PAIR=str(pair)), timeout=300)
            outs[(pair, M)] = np.load(out)
    for M in (3001, 1100):
        assert np.array_equal(outs[(0, M)], outs[(1, M)])


@pytest.mark.parametrize("d,p,N,M,kernel_id", [(10, 1, 4096, 1500, 2), (4, 2, 2000, 1000, 1), (3, 3, 1025, 700, 0)])
def test_int8_large_training_sets(bb, ctx, engine, d, p, N, M, kernel_id):
    """Deep contractions (BASELINE config 5: N = 4096, d = 10) and sizes just past a tile edge: many k-blocks per n-tile, several
    n-tile subsets per query tile, a ragged last tile."""
    rng = np.random.default_rng(N + d)
    X = rng.uniform(-5, 5, (d, N)); Y = rng.standard_normal((p, N))
    lam = rng.uniform(1.5, 4, (d, p)); alpha = rng.uniform(0.5, 2, p); sigma = 0.05 * alpha
    Xq = rng.uniform(-5, 5, (d, M))
    Xq[:, :32] = X[:, :32]
    model = bb.GaussianProcess(kernel_id, lam, alpha, sigma)
    post = bb.model_posterior(model, X, Y, ctx=ctx)
    ofit = orc.gp_fit(kernel_id, X, Y, lam, alpha, sigma)
    omu, ovar = orc.gp_predict(ofit, Xq)
    engine("int8", 7)
    mu, var = post.mean_and_var(Xq)
    engine("int8", 0)                       # the default: 6 slices of K* against 7 of alpha^2 L^-1
    _, var76 = post.mean_and_var(Xq)
    engine("fp64")
    _, var64 = post.mean_and_var(Xq)
    kxx = alpha ** 2
    assert _var_err(var, ovar, kxx) < TOL and _var_err(var, var64, kxx) < TOL
    assert _var_err(var, ovar, kxx) < 4 * _var_err(var64, ovar, kxx) + 1e-12
    assert _var_err(var76, ovar, kxx) < TOL and _var_err(var76, ovar, kxx) < 10 * _var_err(var64, ovar, kxx) + 5e-12
    # the weights a = K^-1 y come from the stepwise triangular solves (64 block steps per direction at N = 4096, many CTAs per step):
    # a and the mean against the oracle at these sizes too
    ag = post.factors()[2]
    assert np.max(np.abs(ag - ofit.a)) / np.max(np.abs(ofit.a)) < 1e-10
    # mu = k* . a is a cancelling sum when the targets are noise (Y ~ N(0, 1) here: |a| ~ |y| / sigma^2 while |mu| = O(1)), so the
    # bound is the one the agreement of `a` implies: 1e-10 of max|a| on every term of the dot product, on top of 1e-10 relative
    for j in range(p):
        Ks = alpha[j] ** 2 * orc.cross_kernel(kernel_id, Xq, X, lam[:, j])            # M x N
        bound = TOL * (np.abs(omu[j]) + 1e-3 * np.max(np.abs(omu))) + TOL * np.max(np.abs(ofit.a[:, j])) * np.abs(Ks).sum(axis=1)
        assert np.all(np.abs(mu[j] - omu[j]) <= bound), float(np.max(np.abs(mu[j] - omu[j]) / bound))
