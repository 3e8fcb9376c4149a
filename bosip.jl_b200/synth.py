"""Seeded synthetic problems of SURVEY.md section 8(d) -- identical inputs for the CUDA path, the oracle and the
CPU baseline.  Shapes follow BASELINE.json `configs` (C1..C5).  Pure NumPy, no GPU needed.

Domain [-5,5]^d and prior/likelihood settings follow /root/reference/examples/simple/toy_problem.jl:15-40."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SE, MATERN32, MATERN52 = 0, 1, 2
UNIFORM, NORMAL, TRUNCNORMAL = 0, 1, 2
BASE_SEED = 20261019

# name -> (config_id, d, p, N, kernel)
CONFIGS = {
    "C1": (1, 2, 1, 20, MATERN52),
    "C2": (2, 2, 2, 200, MATERN52),
    "C3": (3, 5, 4, 1000, MATERN52),
    "C5": (5, 10, 1, 4096, MATERN52),
}


@dataclass
class SynthProblem:
    name: str
    kernel_id: int
    X: np.ndarray        # d x N  training inputs   (Julia layout: column = point)
    Y: np.ndarray        # p x N  simulator outputs
    lam: np.ndarray      # d x p  length-scales
    alpha: np.ndarray    # p      amplitudes
    sigma: np.ndarray    # p      noise std
    z_obs: np.ndarray    # p
    std_obs: np.ndarray  # p
    prior_kind: np.ndarray    # d
    prior_params: np.ndarray  # d x 4  (m, s, a, b) | (a, b, -, -) | (m, s, -, -)
    lb: np.ndarray       # d
    ub: np.ndarray       # d
    seed: int

    @property
    def d(self): return self.X.shape[0]
    @property
    def p(self): return self.Y.shape[0]
    @property
    def N(self): return self.X.shape[1]


def simulator(omega: np.ndarray, X: np.ndarray) -> np.ndarray:
    """y_j(x) = sum_k sin(omega_jk x_k) + 0.1 prod_k x_k/5 ;  X d x M -> p x M."""
    return np.sin(omega[:, :, None] * X[None, :, :]).sum(axis=1) + 0.1 * np.prod(X / 5.0, axis=0)[None, :]


def make_problem(name: str = "C3", *, d=None, p=None, N=None, kernel_id=None, seed=None,
                 noise_frac: float = 0.05) -> SynthProblem:
    cid, d0, p0, N0, k0 = CONFIGS.get(name, (0, 2, 1, 20, MATERN52))
    d = d or d0; p = p or p0; N = N or N0
    kernel_id = k0 if kernel_id is None else kernel_id
    seed = BASE_SEED + cid if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))
    lb, ub = -5.0 * np.ones(d), 5.0 * np.ones(d)
    X = rng.uniform(lb[:, None], ub[:, None], size=(d, N))
    omega = rng.uniform(0.3, 1.5, size=(p, d))
    Y = simulator(omega, X)
    lam = rng.uniform(1.0, 4.0, size=(d, p))
    alpha = Y.std(axis=1) if N > 1 else np.ones(p)
    sigma = noise_frac * alpha
    x_true = rng.uniform(-2.0, 2.0, size=(d, 1))
    std_obs = 0.2 * np.ones(p)
    z_obs = simulator(omega, x_true)[:, 0] + std_obs * rng.standard_normal(p)
    prior_kind = np.full(d, TRUNCNORMAL, dtype=np.int32)
    prior_params = np.tile(np.array([0.0, 5.0 / 3.0, -5.0, 5.0]), (d, 1))
    return SynthProblem(name, kernel_id, X, Y, lam, alpha, sigma, z_obs, std_obs, prior_kind, prior_params, lb, ub, seed)


def grid_queries(lb, ub, n_per_dim: int) -> np.ndarray:
    """Full tensor grid (first coordinate fastest), d x n^d -- C1: 101x101 step 0.1 (examples/simple/main.jl:69)."""
    axes = [np.linspace(l, u, n_per_dim) for l, u in zip(lb, ub)]
    mesh = np.meshgrid(*axes, indexing="ij")
    return np.stack([m.reshape(-1, order="F") for m in mesh], axis=0)


# ---------------------------------------------------------------------------------------------
# Philox4x32-10 keyed by the GLOBAL candidate index, so that candidate i is the same point whichever GPU
# (or the host) generates it.  Mirrors bosip.jl_b200/csrc/philox.cuh bit for bit.
# ---------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32 after 10 rounds."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & _MASK, lo0]
        k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def philox_uniform_candidates(seed: int, start: int, count: int, lb, ub) -> np.ndarray:
    """Candidates start..start+count-1, U(lb, ub)^d, d x count.  Coordinate k of candidate i uses 64 bits:
    block b = k // 2 -> philox(ctr=(i_lo, i_hi, b, 0), key=(seed_lo, seed_hi)), words (2(k%2), 2(k%2)+1);
    u = ((hi << 21) | (lo >> 11)) * 2^-53 in [0,1);  x = lb + (ub - lb) * u."""
    lb = np.asarray(lb, float); ub = np.asarray(ub, float)
    d = lb.shape[0]
    idx = np.arange(start, start + count, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    out = np.empty((d, count))
    for b in range((d + 1) // 2):
        ctr = np.stack([(idx & _MASK).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                        np.full(count, b, np.uint32), np.zeros(count, np.uint32)], axis=-1)
        r = philox4x32(ctr, key).astype(np.uint64)
        for h in range(2):
            k = 2 * b + h
            if k >= d:
                break
            lo, hi = r[:, 2 * h], r[:, 2 * h + 1]
            bits = (hi << np.uint64(21)) | (lo >> np.uint64(11))
            u = bits.astype(np.float64) * (1.0 / 9007199254740992.0)
            out[k] = lb[k] + (ub[k] - lb[k]) * u
    return out
