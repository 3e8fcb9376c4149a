"""The reference's examples/simple toy inverse problem (BASELINE.json config C1) driven through bosip_b200.

Mirrors /root/reference/examples/simple/{main.jl,toy_problem.jl}: 2 parameters in [-5,5]^2, simulator y = x1 * x2, observation
z_obs = 1 with std_obs = 0.2 (toy_problem.jl:13-40), truncated-Normal prior, Matern-5/2 GP (main.jl:36), 3 initial simulations ->
DataLimit(20) (main.jl:21,63), MaxVar acquisition, then approx_posterior / posterior_mean on the 101 x 101 plotting grid (main.jl:69).
Hyper-parameters are fixed here (BOSS's MAP fit is outside this repository's scope).

    python examples/simple.py          # needs a B200
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bosip_b200 as bb  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    lb, ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
    f = lambda x: np.array([x[0] * x[1]])                                    # toy_problem.jl:21
    X0 = rng.uniform(lb[:, None], ub[:, None], size=(2, 3))                   # 3 initial simulations
    Y0 = np.stack([f(X0[:, i]) for i in range(3)], axis=1)
    model = bb.GaussianProcess(bb.MATERN52, lengthscales=np.array([[2.0], [2.0]]), amplitudes=np.array([6.0]), noise_std=np.array([1e-3]))
    like = bb.NormalLikelihood(z_obs=[1.0], std_obs=[0.2])
    prior = bb.ProductPrior([bb.TRUNCNORMAL] * 2, [[0.0, 5.0 / 3.0, -5.0, 5.0]] * 2)
    problem = bb.BosipProblem(X0, Y0, model, like, prior, (lb, ub))
    am = bb.B200CandidateAM(bb.CandidateSource.grid(lb, ub, 401))              # 160 801 candidates per iteration
    bb.bosip(problem, f, acquisition=bb.LogMaxVar(), acq_maximizer=am, term_cond=bb.DataLimit(20), info=True)
    grid = bb.synth.grid_queries(lb, ub, 101)                                  # main.jl:69: step 0.1
    post = bb.posterior_mean(problem, normalize=True, xs=prior.rand(10_000, rng))(grid).reshape(101, 101, order="F")
    i, j = np.unravel_index(np.argmax(post), post.shape)
    print(f"posterior-mean mode at x = ({-5 + 0.1 * i:.1f}, {-5 + 0.1 * j:.1f}); x1*x2 there = {(-5 + 0.1 * i) * (-5 + 0.1 * j):.2f} (z_obs = 1)")
    m = bb.compute_marginals_int(problem, grid_size=50, lhc_grid_size=200, rng=rng)
    print("marginal of x1 integrates to", float((m["marginals"][(0, 0)]["ys"].sum() - 0.5 * (m["marginals"][(0, 0)]["ys"][[0, -1]].sum())) * 10 / 49))


if __name__ == "__main__":
    main()
