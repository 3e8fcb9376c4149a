"""GPU: size-independent properties at BASELINE.json's full per-GPU size (C3: 1e8 candidates / 8 GPUs = 1.25e7 per GPU),
where the oracle is too slow: the sharded argmax composes, chunking and GPU count are invisible, and a sampled subset
still matches the oracle."""
import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu


def test_c3_full_per_gpu_share(bb, ctx):
    pr = bb.synth.make_problem("C3")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs); prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)
    M = 12_500_000
    src = bb.CandidateSource.philox(20261019 + 3, pr.lb, pr.ub, 10 ** 8)
    # rank 3 of 8 owns candidates [3.75e7, 5e7)
    start, count = bb.parallel.shard_range(10 ** 8, 3, 8)
    assert count == M
    whole = bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src, start, count)
    assert start <= whole[0] < start + count and np.isfinite(whole[1])
    # split the same range in 5 uneven pieces: the combined winner is identical (index, value, x)
    cuts = [0, 1_000_003, 4_999_999, 5_000_000, 11_111_111, M]
    parts = [bb.acq_argmax(post, like, prior, bb.LogMaxVar(), src, start + a, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    w = bb.parallel.combine_argmax(np.array([q[1] for q in parts]), np.array([q[0] for q in parts]))
    assert parts[w][0] == whole[0] and parts[w][1] == whole[1]
    np.testing.assert_array_equal(parts[w][2], whole[2])
    # the winner is the same point on host and device and re-evaluates to the same value through the plain closure
    xw = bb.synth.philox_uniform_candidates(20261019 + 3, whole[0], 1, pr.lb, pr.ub)
    np.testing.assert_array_equal(xw[:, 0], whole[2])
    v = bb.posterior_eval(post, like, prior, xw, 4)["log_post_var"][0]
    assert v == whole[1]
    # ... and to the oracle's value, and it dominates an oracle-evaluated sample of the range
    ofit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
    ol = orc.NormalLikelihood(pr.z_obs, pr.std_obs); op = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    assert orc.posterior_eval(ofit, ol, op, xw)["log_post_var"][0] == pytest.approx(whole[1], rel=1e-10)
    sample = bb.synth.philox_uniform_candidates(20261019 + 3, start + 777_777, 20_000, pr.lb, pr.ub)
    assert np.max(orc.posterior_eval(ofit, ol, op, sample)["log_post_var"]) <= whole[1] * (1 - 1e-12) + 1e-9
