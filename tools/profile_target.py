"""Small ncu target: one chunk-sized sweep of the C3 hot path (device-resident inputs)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bosip_b200 as bb

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = bb.Context(0)
if os.environ.get("BOSIP_ENGINE", "fp64") == "int8":
    ctx.set_variance_engine("int8", int(os.environ.get("BOSIP_SLICES", "0")))
pr = bb.synth.make_problem(name)
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
Xd = torch.empty((M, pr.d), dtype=torch.float64, device="cuda").t()
bb.candidates(ctx, bb.CandidateSource.philox(1, pr.lb, pr.ub, M), 0, M, device_out=Xd)
for _ in range(reps):
    r = bb.posterior_eval(post, like, prior, Xd, 7, acq=bb.MaxVar())
print("ok", r["best_idx"], r["best_val"])
