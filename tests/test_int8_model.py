"""CPU: the byte-slicing arithmetic of the INT8 variance engine (tests/int8_model.py, an integer restatement of
csrc/vargemm_i8.cu) against an exact evaluation of q = ||alpha^2 L^{-1} k*||^2 from the same Float64 inputs."""
import os
import sys

import mpmath as mp
import numpy as np
import pytest
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.dirname(__file__))
import int8_model as im  # noqa: E402
from oracle import bosip_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def case():
    rng = np.random.default_rng(5)
    d, N, M = 3, 150, 24
    X = rng.uniform(-5, 5, (d, N)); lam = rng.uniform(1, 4, (d, 1)); alpha = np.array([1.7]); sigma = 0.05 * alpha
    fit = orc.gp_fit(orc.MATERN52, X, np.sin(X).sum(0)[None, :], lam, alpha, sigma)
    B = np.tril(alpha[0] ** 2 * solve_triangular(fit.L[0], np.eye(N), lower=True))      # alpha^2 L^{-1}
    Xq = rng.uniform(-5, 5, (d, M)); Xq[:, :6] = X[:, :6] + 1e-6                       # near training points: alpha^2 - q cancels most
    Ks = orc.cross_kernel(orc.MATERN52, Xq, X, lam[:, 0])                               # unit-scale kernel values in [0, 1]
    mp.mp.dps = 50
    exact = []
    for i in range(M):
        s = mp.mpf(0)
        for n in range(N):
            v = mp.fsum(mp.mpf(float(B[n, k])) * mp.mpf(float(Ks[i, k])) for k in range(n + 1))
            s += v * v
        exact.append(s)
    return Ks, B, alpha[0], exact


def _err(q, exact, a2):
    return max(float(abs(mp.mpf(float(q[i])) - exact[i])) for i in range(len(exact))) / a2


@pytest.mark.parametrize("S,bound", [(6, 3e-12), (7, 2e-14), (8, 5e-15)])
def test_slicing_scheme_error(case, S, bound):
    """Dropping the slice pairs with a + b >= S leaves an error of ~2^-8S of the row scale; at S = 7 (the default) the result is
    within a few Float64 roundings of the exact value, at S = 8 it is as good as a Float64 evaluation, S = 6 is not enough for
    the 1e-10 gate on small variances."""
    Ks, B, alpha, exact = case
    q, x, (IA, IB, f) = im.engine_q(Ks, B, S)           # asserts inside: digits reproduce IB, |D_g| < 2^31, |acc| < 2^63
    assert _err(q, exact, alpha ** 2) < bound
    q64 = ((Ks @ B.T) ** 2).sum(1)                        # plain Float64 evaluation of the same expression
    e64 = _err(q64, exact, alpha ** 2)
    assert e64 < 3e-15
    if S == 7:
        assert _err(q, exact, alpha ** 2) < 8 * e64 + 1e-15
    if S == 8:
        assert _err(q, exact, alpha ** 2) < 2 * e64 + 1e-16


def test_default_scheme_six_kstar_slices(case):
    """The library's default: 7 slices of alpha^2 L^-1, 6 of K* (27 slice pairs).  Leaving the pair (lowest byte of K*) x (highest byte
    of L^-1) out costs a few times the slicing error of the 7 x 7 scheme and stays two orders of magnitude below what 6 x 6 gives."""
    Ks, B, alpha, exact = case
    q76, _, _ = im.engine_q(Ks, B, 7, 6)
    q77, _, _ = im.engine_q(Ks, B, 7)
    q66, _, _ = im.engine_q(Ks, B, 6)
    e76, e77, e66 = (_err(q, exact, alpha ** 2) for q in (q76, q77, q66))
    assert e76 < 6e-14                      # ~ 40 Float64 roundings of alpha^2; the parity gate on sigma^2 >= 6e-4 alpha^2 needs 6e-14
    assert e76 < 20 * e77 + 1e-15 and e76 < e66 / 20


def test_operand_ranges(case):
    Ks, B, alpha, _ = case
    for S in (6, 7, 8):
        f = im.row_exponents(B)
        A, IA = im.slice_a(Ks, S)
        Bp, IB = im.slice_b(B, f, S)
        assert int(IA.max()) <= 1 << (8 * S - 1) and all(0 <= a.min() and a.max() <= 255 for a in A)     # u8 operand
        assert int(np.abs(IB).max()) <= 1 << (8 * S - 2) and all(-128 <= b.min() and b.max() <= 127 for b in Bp)   # s8 operand
        assert np.all(np.max(np.abs(B), axis=1) <= np.ldexp(1.0, f)) and np.all(np.max(np.abs(B), axis=1) > np.ldexp(1.0, f - 1))
