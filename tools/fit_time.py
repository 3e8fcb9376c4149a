"""Fit (GP conditioning) timing on one B200: python tools/fit_time.py [N p d]...  (BOSIP_B200_FIT_TRACE=1 prints the stages)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bosip_b200 as bb
ctx = bb.Context(0)
shapes = [(4096, 4, 10), (4096, 1, 10), (1000, 4, 5)] if len(sys.argv) < 4 else [tuple(int(v) for v in sys.argv[i:i + 3]) for i in range(1, len(sys.argv) - 2, 3)]
for (N, p, d) in shapes:
    pr = bb.synth.make_problem("X", d=d, p=p, N=N, seed=5)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx); t1 = time.perf_counter(); post.free(); t2 = time.perf_counter()
        ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    print(json.dumps({"fit": {"N": N, "p": p, "d": d}, "ms_fit_and_free": ts}), flush=True)
