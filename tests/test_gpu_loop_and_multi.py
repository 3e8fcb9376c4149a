"""GPU: the BO loop mirror (bosip!) on the reference's examples/simple toy problem, and the NCCL exchange steps on 2 GPUs
(skipped on a single-GPU box; the same helpers run over gloo in tests/test_host_logic.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_bo_loop_simple_toy_problem(bb, ctx):
    """examples/simple (config C1): 3 initial simulations -> DataLimit(12); the approximate posterior concentrates on x1*x2 = 1."""
    rng = np.random.default_rng(0)
    lb, ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
    f = lambda x: np.array([x[0] * x[1]])
    X0 = rng.uniform(lb[:, None], ub[:, None], size=(2, 3))
    Y0 = np.stack([f(X0[:, i]) for i in range(3)], axis=1)
    model = bb.GaussianProcess(bb.MATERN52, np.array([[2.0], [2.0]]), np.array([6.0]), np.array([1e-3]))
    like = bb.NormalLikelihood([1.0], [0.2]); prior = bb.ProductPrior([bb.TRUNCNORMAL] * 2, [[0.0, 5.0 / 3.0, -5.0, 5.0]] * 2)
    prob = bb.BosipProblem(X0, Y0, model, like, prior, (lb, ub), ctx)
    am = bb.B200CandidateAM(bb.CandidateSource.grid(lb, ub, 201))
    picked = []
    bb.bosip(prob, f, acquisition=bb.LogMaxVar(), acq_maximizer=am, term_cond=bb.DataLimit(12), callback=lambda p, it, x, v: picked.append(x.copy()))
    assert prob.X.shape == (2, 12) and prob.Y.shape == (1, 12) and len(picked) == 9
    assert len({tuple(np.round(x, 6)) for x in picked}) == 9                 # MaxVar never re-queries a simulated point
    grid = bb.synth.grid_queries(lb, ub, 101)
    lp = bb.log_posterior_mean(prob)(grid)
    xm = grid[:, int(np.argmax(lp))]
    assert abs(xm[0] * xm[1] - 1.0) < 0.5                                     # the mode sits on the hyperbola x1*x2 = z_obs


WORKER = r'''
import os, sys, math
import numpy as np, torch
sys.path.insert(0, os.environ["BOSIP_ROOT"])
import torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import bosip_b200 as bb
ctx = bb.Context(rank)
pr = bb.synth.make_problem("C2")
model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
prob = bb.BosipProblem(pr.X, pr.Y, model, like, prior, (pr.lb, pr.ub), ctx)
src = bb.CandidateSource.philox(5, pr.lb, pr.ub, 200001)
x, val = bb.B200CandidateAM(src).maximize_acquisition(prob, bb.LogMaxVar())     # sharded + NCCL all-gather
import torch.distributed as d2
d2.destroy_process_group() if False else None
# single-process reference on this rank (no sharding)
full = bb.acq_argmax(prob.model_posterior(), like, prior, bb.LogMaxVar(), src, 0, 200001)
assert full[1] == val and np.array_equal(full[2], x), (full, val, x)
rng = np.random.default_rng(0)
A = rng.standard_normal((3, 2000)); B = 0.2 + rng.standard_normal((3, 1500)); k = bb.Kernel(bb.SE, np.ones(3))
v = bb.calc_mmd(k, A, B, ctx)                                                      # tile grid dealt across ranks + all-reduce
G = 40001; lw = rng.normal(0, .1, G); a = rng.normal(-3, 2, G); t = a + rng.normal(0, .5, G)
s, c = bb.parallel.shard_range(G)
tv = bb.calc_tv(lw[s:s + c], a[s:s + c], t[s:s + c], ctx)                           # each rank passes its slice
from oracle import bosip_oracle as orc
assert abs(v - orc.calc_mmd(orc.SE, np.ones(3), A, B)[0]) < 1e-12 * max(1.0, abs(v)) + 1e-13
assert abs(tv - orc.calc_tv(lw, a, t)) < 1e-10 * tv
dist.barrier(); dist.destroy_process_group()
print("MULTI_OK", rank)
'''


def test_nccl_exchange_steps_on_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, BOSIP_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"MULTI_OK {r}" in o, o[-3000:]
