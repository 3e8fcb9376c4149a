import sys, os, time, json
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bosip_b200.py") else os.getcwd())
import numpy as np, torch
import bosip_b200 as bb
ctx = bb.Context(0)
N = int(sys.argv[1]); M = 148 * 128 * 2
pr = bb.synth.make_problem("X", d=10, p=1, N=N, seed=5)
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
Xd = torch.empty((M, pr.d), dtype=torch.float64, device="cuda").t()
bb.candidates(ctx, bb.CandidateSource.philox(99, pr.lb, pr.ub, M), 0, M, device_out=Xd)
bb.posterior_eval(post, like, prior, Xd, 7)
ctx.timing_reset(True)
bb.posterior_eval(post, like, prior, Xd, 7); bb.posterior_eval(post, like, prior, Xd, 7)
torch.cuda.synchronize()
tm = ctx.timing_get()
print(json.dumps({"N": N, "nsub_env": os.environ.get("BOSIP_I8_NSUB"), "ms": {k: v["ms"] for k, v in tm.items()}, "launches": tm["var_gemm"]["launches"]}))
