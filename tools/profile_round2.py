"""ncu targets of round 2 (one B200):  python tools/profile_round2.py c3 | tv | fit
   c3 : two argmax sweeps of one C3 chunk with device-generated candidates (candidates_kernel, kstar_kernel, vargemm_i8_kernel, epilogue_kernel ...)
   tv : three calc_tv calls at G = 1e8 (tv_stage1_kernel / tv_stage2_kernel)
   fit: two GP fits at N = 4096, p = 1 (gemm_dmma_kernel, potrf_diag_kernel, trsv_*_step)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bosip_b200 as bb

what = sys.argv[1] if len(sys.argv) > 1 else "c3"
ctx = bb.Context(0)
if what == "c3":
    pr = bb.synth.make_problem("C3")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    M = 148 * 128 * 4
    src = bb.CandidateSource.philox(1, pr.lb, pr.ub, M)
    for _ in range(2):
        r = bb.acq_argmax(post, like, prior, bb.MaxVar(), src)
    print("ok", r[0], r[1])
elif what == "tv":
    G = 10 ** 8
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    lw = 0.1 * torch.randn(G, dtype=torch.float64, device="cuda", generator=g)
    a = -3.0 + 2.0 * torch.randn(G, dtype=torch.float64, device="cuda", generator=g)
    t = a + 0.5 * torch.randn(G, dtype=torch.float64, device="cuda", generator=g)
    for _ in range(3):
        v = bb.calc_tv(lw, a, t, ctx)
    print("ok", v)
else:
    pr = bb.synth.make_problem("X", d=10, p=1, N=4096, seed=5)
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    for _ in range(2):
        post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx); post.free()
    print("ok")
