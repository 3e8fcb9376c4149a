"""GPU: MMD (src/metrics/mmd.jl:23-28) and TV (src/metrics/tv.jl:59-73) through the C ABI vs oracle and mpmath goldens."""
import json
import math
import os

import numpy as np
import pytest

from oracle import bosip_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics_small.json")))


def test_mmd_goldens(bb, ctx):
    for c in GOLD["mmd"]:
        v = bb.calc_mmd(bb.Kernel(c["kernel_id"], np.array(c["lam"])), np.array(c["A"]), np.array(c["B"]), ctx)
        assert v == pytest.approx(c["mmd2"], rel=1e-10, abs=1e-15)


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
@pytest.mark.parametrize("d,n,m", [(1, 1, 1), (2, 127, 129), (5, 2048, 2049), (5, 5000, 3000), (10, 300, 4100)])
def test_mmd_vs_oracle(bb, ctx, kernel_id, d, n, m):
    rng = np.random.default_rng(100 * d + n)
    A = rng.standard_normal((d, n)); B = 0.1 + 1.1 * rng.standard_normal((d, m)); lam = rng.uniform(0.5, 2.0, d)
    sums = np.zeros(3)
    lib = ctx.lib
    import ctypes as C
    Af, Bf = np.asfortranarray(A), np.asfortranarray(B)
    ctx._check(lib.bosip_b200_mmd_sums(ctx.h, kernel_id, d, lam.ctypes.data_as(bb._lib.c_double_p), Af.ctypes.data, n, Bf.ctypes.data, m,
                                       0, 0, 1, sums.ctypes.data_as(bb._lib.c_double_p)))
    o = orc.calc_mmd(kernel_id, lam, A, B)
    np.testing.assert_allclose(sums, o[1:], rtol=1e-10)                       # the three V-statistic sums, diagonals included
    assert bb.calc_mmd(bb.Kernel(kernel_id, lam), A, B, ctx) == pytest.approx(o[0], rel=1e-9, abs=1e-13)
    # the rank-dealt tile grid composes exactly to the whole (what the NCCL all-reduce sums)
    parts = np.zeros((4, 3))
    for r in range(4):
        ctx._check(lib.bosip_b200_mmd_sums(ctx.h, kernel_id, d, lam.ctypes.data_as(bb._lib.c_double_p), Af.ctypes.data, n, Bf.ctypes.data,
                                           m, 0, r, 4, parts[r].ctypes.data_as(bb._lib.c_double_p)))
    np.testing.assert_allclose(parts.sum(0), sums, rtol=1e-12)


def test_mmd_properties(bb, ctx):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((5, 1500)); k = bb.Kernel(bb.SE, np.ones(5))
    assert abs(bb.calc_mmd(k, A, A, ctx)) < 1e-13                              # identical sample sets
    B = A + 0.5
    assert bb.calc_mmd(k, A, B, ctx) == pytest.approx(bb.calc_mmd(k, B, A, ctx), rel=1e-12)   # symmetry
    import torch
    Ad = torch.as_tensor(np.ascontiguousarray(A.T), device="cuda").t(); Bd = torch.as_tensor(np.ascontiguousarray(B.T), device="cuda").t()
    assert bb.calc_mmd(k, Ad, Bd, ctx) == bb.calc_mmd(k, A, B, ctx)            # device-resident == host
    assert bb.calculate_metric(bb.MMDMetric(k), A, B, ctx) == bb.calc_mmd(k, A, B, ctx)


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
def test_mmd_distance_forms(bb, ctx, kernel_id):
    """The pair kernel forms squared distances as |p|^2 + |q|^2 - 2 p.q when all centred norms are small and as
    sum (p - q)^2 otherwise (csrc/metrics.cu); both agree with the oracle, and a common translation changes nothing."""
    rng = np.random.default_rng(17 + kernel_id)
    d, n, m = 4, 1500, 1100
    lam = rng.uniform(0.8, 1.5, d); k = bb.Kernel(kernel_id, lam)
    A = 0.4 * rng.standard_normal((d, n)); B = 0.1 + 0.5 * rng.standard_normal((d, m))          # compact: expanded form
    v = bb.calc_mmd(k, A, B, ctx)
    assert v == pytest.approx(orc.calc_mmd(kernel_id, lam, A, B)[0], rel=1e-10, abs=1e-14)
    shift = np.array([1e3, -2e3, 5e2, 7e3])[:, None]                                               # centring removes the offset
    assert bb.calc_mmd(k, A + shift, B + shift, ctx) == pytest.approx(v, rel=1e-8, abs=1e-13)
    Aw = 30.0 * rng.standard_normal((d, n)); Bw = 3.0 + 33.0 * rng.standard_normal((d, m))         # many length-scales wide: difference form
    assert bb.calc_mmd(k, Aw, Bw, ctx) == pytest.approx(orc.calc_mmd(kernel_id, lam, Aw, Bw)[0], rel=1e-10, abs=1e-14)
    Bfar = B + 1e4                                                                                 # the two sets far apart: cross term underflows to 0
    assert bb.calc_mmd(k, A, Bfar, ctx) == pytest.approx(orc.calc_mmd(kernel_id, lam, A, Bfar)[0], rel=1e-10)
    An = A.copy(); An[1, 7] = np.nan
    assert np.isnan(bb.calc_mmd(k, An, B, ctx))


def test_tv_goldens_and_oracle(bb, ctx):
    for c in GOLD["tv"]:
        assert bb.calc_tv(c["log_ws"], c["a"], c["t"], ctx) == pytest.approx(c["tv"], rel=1e-10, abs=1e-16)
    rng = np.random.default_rng(8)
    for G in (1, 2, 255, 256, 257, 100_003, 1_000_000):
        lw = rng.normal(0, 0.3, G); a = rng.normal(-5, 4, G); t = a + rng.normal(0, 1, G)
        assert bb.calc_tv(lw, a, t, ctx) == pytest.approx(orc.calc_tv(lw, a, t), rel=1e-10, abs=1e-16)


def test_tv_edge_cases(bb, ctx):  # tv.jl:64-66, utils.jl:9
    lw = np.zeros(8); a = np.linspace(-3, 0, 8)
    assert bb.calc_tv(lw, a, a, ctx) == 0.0
    t = np.full(8, -np.inf); t[0] = 0.0
    b = np.full(8, -np.inf); b[7] = 0.0
    assert bb.calc_tv(lw, b, t, ctx) == pytest.approx(1.0, rel=1e-14)
    assert bb.calc_tv(lw, np.full(8, -np.inf), np.full(8, -np.inf), ctx) == 0.0
    # partially -Inf inputs (points outside the prior support) agree with the oracle
    rng = np.random.default_rng(2)
    G = 5000; lw = rng.normal(0, 0.1, G); a = rng.normal(-2, 2, G); t = rng.normal(-2, 2, G)
    a[::7] = -np.inf; t[::5] = -np.inf
    assert bb.calc_tv(lw, a, t, ctx) == pytest.approx(orc.calc_tv(lw, a, t), rel=1e-10)


def test_tv_stages_compose_across_shards(bb, ctx):
    """The two-stage (max, sum) exchange over 3 'ranks' reproduces the single-GPU value."""
    import ctypes as C
    rng = np.random.default_rng(4)
    G = 30_001; lw = rng.normal(0, 0.2, G); a = rng.normal(-4, 3, G); t = a + rng.normal(0, 0.7, G)
    P = bb.parallel
    cuts = [P.shard_range(G, r, 3) for r in range(3)]
    dp = bb._lib.c_double_p
    s1 = []
    for s, c in cuts:
        o = np.zeros(4)
        ctx._check(ctx.lib.bosip_b200_tv_stage1(ctx.h, lw[s:s + c].ctypes.data, a[s:s + c].ctypes.data, t[s:s + c].ctypes.data, c, 0, o.ctypes.data_as(dp)))
        s1.append(o)
    ma = sa = mt = st = None
    Ma, Sa, Mt, St = -math.inf, 0.0, -math.inf, 0.0
    for o in s1:
        Ma, Sa = P.lse_merge(Ma, Sa, o[0], o[1]); Mt, St = P.lse_merge(Mt, St, o[2], o[3])
    eva, evt = Ma + math.log(Sa) - math.log(G), Mt + math.log(St) - math.log(G)
    Md, Sd = -math.inf, 0.0
    for s, c in cuts:
        o = np.zeros(2)
        ctx._check(ctx.lib.bosip_b200_tv_stage2(ctx.h, lw[s:s + c].ctypes.data, a[s:s + c].ctypes.data, t[s:s + c].ctypes.data, c, 0, eva, evt, o.ctypes.data_as(dp)))
        Md, Sd = P.lse_merge(Md, Sd, o[0], o[1])
    tv = 0.5 * math.exp(Md + math.log(Sd) - math.log(G))
    assert tv == pytest.approx(bb.calc_tv(lw, a, t, ctx), rel=1e-13)
    assert tv == pytest.approx(orc.calc_tv(lw, a, t), rel=1e-10)


def test_tv_metric_object(bb, ctx):
    rng = np.random.default_rng(6)
    grid = rng.uniform(-5, 5, (2, 4000)); ws = np.full(4000, 100.0)
    true_lp = lambda X: -0.5 * (X ** 2).sum(0)
    approx_lp = lambda X: -0.5 * ((X - 0.3) ** 2).sum(0) / 1.2
    m = bb.TVMetric(grid, ws=ws, true_logpost=true_lp)
    v = bb.calculate_metric(m, true_lp, approx_lp, ctx)
    assert v == pytest.approx(orc.calc_tv(np.log(ws), approx_lp(grid), true_lp(grid)), rel=1e-10)
    assert 0.0 < v < 1.0
