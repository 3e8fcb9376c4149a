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
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
X = bb.synth.philox_uniform_candidates(20261022, 37_500_000, M, pr.lb, pr.ub)
mu, var = post.mean_and_var(X)
omu, ovar = orc.gp_predict(ofit, X)
err = np.abs(var - ovar) / (pr.alpha[:, None] ** 2)
bad = np.argwhere(err > 1e-9)
print("M", M, "max var err/kxx", err.max(), "n bad", len(bad))
for j, i in bad[:20]:
    print("out", j, "pt", i, "tile", i // 128, "row", i % 128, "gpu", var[j, i], "oracle", ovar[j, i])
if len(bad):
    import collections
    print("bad per output", collections.Counter(bad[:, 0]))
    print("bad rows mod 128", sorted(collections.Counter(bad[:, 1] % 128).items())[:40])
    print("bad tiles", sorted(collections.Counter(bad[:, 1] // 128).items())[:40])
