import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bosip_b200 as bb
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 16
ctx = bb.Context(0)
pr = bb.synth.make_problem(name)
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
Xd = torch.empty((M, pr.d), dtype=torch.float64, device="cuda").t()
bb.candidates(ctx, bb.CandidateSource.philox(1, pr.lb, pr.ub, M), 0, M, device_out=Xd)
bb.posterior_eval(post, like, prior, Xd, 7, acq=bb.MaxVar())
ctx.timing_reset(True)
t0 = time.time(); r = bb.posterior_eval(post, like, prior, Xd, 7, acq=bb.MaxVar()); dt = time.time() - t0
tm = ctx.timing_get()
print(json.dumps({"cfg": name, "M": M, "pts_per_s": M / dt, "ms": {k: v["ms"] / max(1, v["launches"]) for k, v in tm.items()}, "launches": tm["var_gemm"]["launches"], "env": os.environ.get("BOSIP_VG_NOIMAX")}))
