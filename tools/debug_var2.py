import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bosip_b200 as bb
from oracle import bosip_oracle as orc
ctx = bb.Context(0)
pr = bb.synth.make_problem("C3")
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
for M in [int(a) for a in sys.argv[1:]]:
    X = bb.synth.philox_uniform_candidates(20261022, 0, M, pr.lb, pr.ub)
    mu, var = post.mean_and_var(X)
    sel = np.unique(np.concatenate([np.arange(0, min(M, 2000)), np.arange(max(0, M - 2000), M), np.random.default_rng(0).integers(0, M, 3000)]))
    omu, ovar = orc.gp_predict(ofit, X[:, sel])
    err = np.abs(var[:, sel] - ovar) / (pr.alpha[:, None] ** 2)
    emu = np.abs(mu[:, sel] - omu).max()
    print("M", M, "max var err/kxx", err.max(), "frac bad", (err > 1e-9).mean(), "mu err", emu, flush=True)
