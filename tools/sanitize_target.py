"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel family once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bosip_b200 as bb
ctx = bb.Context(0)
pr = bb.synth.make_problem("X", d=3, p=2, N=150, seed=5)
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
rng = np.random.default_rng(0)
Xq = rng.uniform(pr.lb[:, None], pr.ub[:, None], size=(3, 333))
ctx.set_chunk(128)
r = bb.posterior_eval(post, like, prior, Xq, 7, acq=bb.MaxVar())
mu, var = post.mean_and_var(Xq)
src = bb.CandidateSource.philox(1, pr.lb, pr.ub, 1000)
print(bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src)[:2])
prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx)
print(bb.evidence(prob, "mean", xs=prior.rand(500, rng)))
m = bb.compute_marginals_int(prob, grid_size=5, lhc_grid_size=11, rng=rng)
posts = [post, bb.model_posterior(bb.GaussianProcess(pr.kernel_id, 1.2 * pr.lam, pr.alpha, pr.sigma), pr.X, pr.Y, ctx=ctx)]
rb = bb.posterior_eval(posts, like, prior, Xq, 7)
A = rng.standard_normal((3, 300)); B = rng.standard_normal((3, 257))
print(bb.calc_mmd(bb.Kernel(bb.MATERN32, np.ones(3)), A, B, ctx))
G = 5000
print(bb.calc_tv(rng.normal(0, .1, G), rng.normal(-2, 1, G), rng.normal(-2, 1, G), ctx))
print(bb.normal_likelihood_eval(like, mu, var)["log_like_var"][:2])
print("sanitize target ok", r["best_idx"])
