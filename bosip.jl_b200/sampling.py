"""Host-side mirror of the reference's sampler / confidence-set / OptMMD interfaces around the device kernels
(SURVEY.md section 8f ranks 2 and 4).  Names, argument meaning and error behaviour follow the reference:

  AMIS, NormalProposal, AnalyticalFitter, AMISSampler, sample_posterior, resample, sample_posterior_pure
      /root/reference/src/samplers/amis/amis_method.jl:26-134, amis.jl:30-61, proposal_distributions/normal.jl,
      distribution_fitters/analytical.jl, src/sampling.jl:55-83
  find_cutoff, approx_cutoff_area, set_iou, AEConfidence
      /root/reference/src/utils/confidence_set.jl:19-81, src/term_conds/ae_confidence.jl:19-79
  OptMMDMetric
      /root/reference/src/metrics/opt_mmd.jl:17-54

What runs on the GPU: the proposal draws, log_pi(xs) (the fused posterior sweep), the proposal-mixture pdf sums, the log-weights,
exp_weights + weighted mean/covariance, the multinomial down-sampling, calc_mmd, and the confidence-set reductions (weighted
quantile by a device sort + scan, cut-off area, IoU).  What stays on the host: the d x d Cholesky of a proposal covariance and
the length-scale optimiser of OptMMD."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _lib as L
from .api import (BosipB200Error, Context, Kernel, _colmajor, _dptr, _is_torch, _out, approx_posterior, calc_mmd,
                  default_context, posterior_mean)


# ----------------------------------------------------------------------------------------------- proposal distribution
class NormalProposal:
    """NormalProposal(MvNormal(mu, Sigma)) -- proposal_distributions/normal.jl:4-15"""

    def __init__(self, mu, Sigma):
        self.set(mu, Sigma)

    def set(self, mu, Sigma):
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        Sigma = np.asarray(Sigma, dtype=np.float64)
        if Sigma.ndim == 1:
            Sigma = np.diag(Sigma)
        self.Sigma = 0.5 * (Sigma + Sigma.T)
        self.Lc = np.linalg.cholesky(self.Sigma)                    # raises LinAlgError if not positive definite (isposdef)
        self.Linv = np.linalg.inv(self.Lc)
        self.lognorm = -float(np.sum(np.log(np.diag(self.Lc)))) - 0.5 * len(self.mu) * math.log(2.0 * math.pi)

    @property
    def x_dim(self) -> int:
        return len(self.mu)

    def copy(self) -> "NormalProposal":
        return NormalProposal(self.mu.copy(), self.Sigma.copy())

    def rand(self, count: int, seed: int, offset: int = 0, ctx: Optional[Context] = None, device: bool = False):
        """rand(q, count) -> d x count (normal.jl:19-21); draw i depends only on (seed, offset + i)."""
        ctx = ctx or default_context()
        d = self.x_dim
        if device:
            import torch
            out = torch.empty((count, d), dtype=torch.float64, device=f"cuda:{ctx.device}").t()
            ptr, mem = out.data_ptr(), L.DEVICE
        else:
            out = np.empty((d, count), order="F"); ptr, mem = out.ctypes.data, L.HOST
        Lf = np.asfortranarray(self.Lc)
        if device:
            ctx.order_after_torch()
        ctx._check(ctx.lib.bosip_b200_mvnormal_sample(ctx.h, d, _dptr(self.mu), _dptr(Lf), int(seed), int(offset), int(count), mem, ptr))
        return out

    def logpdf(self, X):
        X = np.asarray(X, dtype=np.float64)
        z = self.Linv @ (X.reshape(self.x_dim, -1) - self.mu[:, None])
        return self.lognorm - 0.5 * np.sum(z * z, axis=0)


def mixture_pdf_sum(qs, X, delta=None, ctx: Optional[Context] = None):
    """delta (+)= sum_q pdf(q, x) for every column x of X (amis_method.jl:52,66-68,74).  Returns delta (allocated when None)."""
    ctx = ctx or default_context()
    d = qs[0].x_dim
    kx, px, mem, M = _colmajor(X, d)
    acc = delta is not None
    if delta is None:
        delta, pd = _out((M,), mem, like=kx if mem == L.DEVICE else None)
    else:
        kd, pd, md, _ = _colmajor(delta)
        if md != mem:
            raise ValueError("X and delta must live in the same memory space")
        if kd is not delta:
            raise ValueError("delta must be a contiguous float64 buffer (it is updated in place)")
    mus = np.asfortranarray(np.stack([q.mu for q in qs], axis=1))
    Linvs = np.ascontiguousarray(np.stack([np.asfortranarray(q.Linv).T for q in qs]))   # each block: column-major d x d
    ln = np.ascontiguousarray([q.lognorm for q in qs], dtype=np.float64)
    if mem == L.DEVICE:
        ctx.order_after_torch()
    ctx._check(ctx.lib.bosip_b200_mixture_pdf_sum(ctx.h, d, len(qs), _dptr(mus), _dptr(Linvs), _dptr(ln), px, M, mem, 1 if acc else 0, pd))
    return delta


def amis_log_weights(log_p, delta, t: float, out=None, ctx: Optional[Context] = None):
    """log_Omega = log_P - log(Delta / t)  (amis_method.jl:53,69,75)"""
    ctx = ctx or default_context()
    kp, pp, mem, M = _colmajor(log_p); kd, pd, m2, _ = _colmajor(delta)
    if mem != m2:
        raise ValueError("inputs must live in the same memory space")
    if out is None:
        out, po = _out((M,), mem, like=kp if mem == L.DEVICE else None)
    else:
        ko, po, _, _ = _colmajor(out)
    if mem == L.DEVICE:
        ctx.order_after_torch()
    ctx._check(ctx.lib.bosip_b200_amis_log_weights(ctx.h, pp, pd, float(t), M, mem, po))
    return out


def weighted_moments(X, log_ws, want_ws: bool = False, ctx: Optional[Context] = None):
    """exp_weights + estimate_parameters! sums: returns (mean, cov, V1, V2[, ws]) with w = exp(log_ws - max)."""
    ctx = ctx or default_context()
    kx, px, mem, M = _colmajor(X)
    d = kx.shape[0]
    kw, pw, m2, _ = _colmajor(log_ws)
    if mem != m2:
        raise ValueError("inputs must live in the same memory space")
    mean = np.zeros(d); cov = np.zeros((d, d), order="F"); sums = np.zeros(2)
    ws, pws = (None, None)
    if want_ws:
        ws, pws = _out((M,), mem, like=kx if mem == L.DEVICE else None)
    if mem == L.DEVICE:
        ctx.order_after_torch()
    try:
        ctx._check(ctx.lib.bosip_b200_weighted_moments(ctx.h, d, px, pw, M, mem, _dptr(mean), _dptr(cov), _dptr(sums), pws))
    except BosipB200Error as e:
        if e.code == L.EDOMAIN:
            raise RuntimeError("AMIS: Numerical issues with sample weights.") from e    # amis_method.jl:112-114
        raise
    return (mean, np.ascontiguousarray(cov), sums[0], sums[1], ws) if want_ws else (mean, np.ascontiguousarray(cov), sums[0], sums[1])


def exp_weights(log_ws, ctx: Optional[Context] = None):
    """exp_weights(log_ws) -- amis_method.jl:126-134: exp(log_ws - max) / sum"""
    X = np.zeros((1, len(log_ws))) if not _is_torch(log_ws) else log_ws.reshape(1, -1) * 0.0
    return weighted_moments(X, log_ws, want_ws=True, ctx=ctx)[4]


class AnalyticalFitter:
    """fit_distribution!(::AnalyticalFitter, dist, xs, ws) -> estimate_parameters! (analytical.jl:9-13, normal.jl:75-100).
    Takes the LOG weights of the samples (the normalisation of exp_weights cancels in the estimator)."""

    def fit(self, dist: NormalProposal, xs, log_ws, ctx: Optional[Context] = None) -> NormalProposal:
        mean, cov, V1, V2 = weighted_moments(xs, log_ws, ctx=ctx)
        ok = np.all(np.isfinite(cov))
        if ok:
            try:
                dist.set(mean, cov)
            except np.linalg.LinAlgError:
                ok = False
        # not positive definite / infinite: "Reusing parameters from the previous iteration." (normal.jl:92-95)
        return dist


# ----------------------------------------------------------------------------------------------- AMIS
@dataclass
class AMIS:
    """AMIS(; T, N, init_q); (amis)(log_pi, q, fitter) -> (xs, ws) -- amis_method.jl:18-83.
    `log_pi` maps a d x M matrix (NumPy, or a CUDA tensor when device=True) to M log-density values."""
    T: int = 10
    N: int = 20
    init_q: Optional[NormalProposal] = None

    def __call__(self, log_pi: Callable, q: NormalProposal, fitter: AnalyticalFitter, seed: int = 0, ctx: Optional[Context] = None,
                 device: bool = True, return_trace: bool = False):
        ctx = ctx or default_context()
        d, N, T = q.x_dim, self.N, self.T
        qs = [q.copy() for _ in range(T + 1)]
        if self.init_q is not None:
            qs[0] = self.init_q
        if device:
            import torch
            dev = f"cuda:{ctx.device}"
            xs = torch.empty(((T + 1) * N, d), dtype=torch.float64, device=dev).t()       # d x (T+1)N, column-major
            log_P = torch.empty((T + 1) * N, dtype=torch.float64, device=dev)
            Delta = torch.zeros((T + 1) * N, dtype=torch.float64, device=dev)
            log_O = torch.empty((T + 1) * N, dtype=torch.float64, device=dev)
        else:
            xs = np.empty((d, (T + 1) * N), order="F"); log_P = np.empty((T + 1) * N); Delta = np.zeros((T + 1) * N); log_O = np.empty((T + 1) * N)

        def cols(a, t0, t1):
            return a[:, t0 * N:t1 * N] if getattr(a, "ndim", None) == 2 or (_is_torch(a) and a.dim() == 2) else a[t0 * N:t1 * N]

        def draw(t):
            x = qs[t].rand(N, seed, offset=t * N, ctx=ctx, device=device)
            if device:
                cols(xs, t, t + 1).copy_(x)
            else:
                cols(xs, t, t + 1)[...] = x

        def put(dst, t0, t1, val):
            if device:
                import torch
                cols(dst, t0, t1).copy_(val if _is_torch(val) else torch.as_tensor(np.asarray(val), device=dst.device))
            else:
                cols(dst, t0, t1)[...] = np.asarray(val)

        # t = 0 (amis_method.jl:46-53)
        draw(0)
        put(log_P, 0, 1, log_pi(cols(xs, 0, 1)))
        mixture_pdf_sum([qs[0]], cols(xs, 0, 1), cols(Delta, 0, 1), ctx)
        amis_log_weights(cols(log_P, 0, 1), cols(Delta, 0, 1), 1.0, out=cols(log_O, 0, 1), ctx=ctx)
        for t in range(1, T + 1):
            # fit the next proposal on all samples drawn so far (amis_method.jl:58-59); get_weights -> exp_weights normalises,
            # which the estimator is invariant to, so the log-weights go in directly (errors are raised as the reference does)
            fitter.fit(qs[t], cols(xs, 0, t), cols(log_O, 0, t), ctx)
            draw(t)
            put(log_P, t, t + 1, log_pi(cols(xs, t, t + 1)))
            mixture_pdf_sum(qs[:t + 1], cols(xs, t, t + 1), cols(Delta, t, t + 1), ctx)            # amis_method.jl:66-68
            mixture_pdf_sum([qs[t]], cols(xs, 0, t), cols(Delta, 0, t), ctx)                        # amis_method.jl:73-74
            amis_log_weights(cols(log_P, 0, t + 1), cols(Delta, 0, t + 1), float(t + 1), out=cols(log_O, 0, t + 1), ctx=ctx)
        xs_ = cols(xs, 1, T + 1)
        ws_ = weighted_moments(xs_, cols(log_O, 1, T + 1), want_ws=True, ctx=ctx)[4]                # get_weights(log_Omega, 2:T+1)
        if return_trace:
            return xs_, ws_, {"xs": xs, "log_P": log_P, "Delta": Delta, "log_Omega": log_O, "qs": qs}
        return xs_, ws_


@dataclass
class AMISSampler:
    """AMISSampler(; iters, proposal_fitter) -- amis.jl:30-34.  The Gaussian-mixture initial proposal (approx_by_gauss_mix) is
    outside the evaluator's scope: pass `init_q` explicitly or start from the broad default proposal (amis.jl:46-50)."""
    iters: int
    proposal_fitter: AnalyticalFitter = field(default_factory=AnalyticalFitter)
    init_q: Optional[NormalProposal] = None


def sample_posterior(sampler: AMISSampler, logpost: Callable, bounds, count: int, seed: int = 0, ctx: Optional[Context] = None,
                     device: bool = True):
    """sample_posterior(::AMISSampler, logpost, domain, count) -> (xs, ws) -- amis.jl:36-61"""
    lb, ub = (np.asarray(b, dtype=np.float64) for b in bounds)
    per_iter = int(math.ceil(sampler.iters * count / sampler.iters))
    q = NormalProposal(0.5 * (lb + ub), np.diag((ub - lb) / 5.0))
    amis = AMIS(T=sampler.iters, N=per_iter, init_q=sampler.init_q)
    return amis(logpost, q, sampler.proposal_fitter, seed=seed, ctx=ctx, device=device)


def resample(xs, ws, count: int, seed: int = 0, ctx: Optional[Context] = None, return_index: bool = False):
    """resample(xs, ws, count) -- src/sampling.jl:55-57: `count` columns drawn with replacement, probability proportional to ws."""
    ctx = ctx or default_context()
    kx, px, mem, M = _colmajor(xs)
    d = kx.shape[0]
    kw, pw, m2, _ = _colmajor(ws)
    if mem != m2:
        raise ValueError("inputs must live in the same memory space")
    out, po = _out((d, count), mem, like=kx if mem == L.DEVICE else None)
    idx, pi = None, None
    if return_index:
        if mem == L.DEVICE:
            import torch
            idx = torch.empty(count, dtype=torch.int64, device=kx.device); pi = idx.data_ptr()
        else:
            idx = np.empty(count, dtype=np.int64); pi = idx.ctypes.data
    if mem == L.DEVICE:
        ctx.order_after_torch()
    ctx._check(ctx.lib.bosip_b200_resample(ctx.h, d, px, pw, M, int(seed), int(count), mem, po, pi))
    return (out, idx) if return_index else out


def sample_posterior_pure(sampler: AMISSampler, logpost: Callable, bounds, count: int, supersample_ratio: int = 20, seed: int = 0,
                          ctx: Optional[Context] = None, device: bool = True):
    """sample_posterior_pure(::WeightedSampler, ...) -- src/sampling.jl:77-83"""
    xs, ws = sample_posterior(sampler, logpost, bounds, supersample_ratio * count, seed=seed, ctx=ctx, device=device)
    return resample(xs, ws, count, seed=seed + 1, ctx=ctx)


# ----------------------------------------------------------------------------------------------- confidence sets
def _vec(x):
    """(keepalive, address, mem, n) of a length-n float64 vector (NumPy or torch)."""
    return _colmajor(x if _is_torch(x) else np.asarray(x, dtype=np.float64).reshape(-1))


def weighted_quantile(vals, ws, p: float, ctx: Optional[Context] = None) -> float:
    """quantile(vals, Distributions.weights(ws), p) as StatsBase.jl defines it for non-frequency weights (zero weights dropped,
    values sorted, h = p (W - w_1) + w_1 on the cumulative weights, linear interpolation) -- on the device: bitonic sort of
    (value, index) pairs, fixed-order prefix sum of the weights, one interpolation (csrc/confset.cu).  ws=None: unit weights."""
    ctx = ctx or default_context()
    kv, pv, mem, n = _vec(vals)
    if n == 0:
        raise ValueError("quantile of an empty array is undefined")
    kw = pw = None
    if ws is not None:
        kw, pw, m2, n2 = _vec(ws)
        if m2 != mem or n2 != n:
            raise ValueError("vals and ws must have the same length and live in the same memory space")
    if mem == L.DEVICE:
        ctx.order_after_torch()
    out = C.c_double()
    ctx._check(ctx.lib.bosip_b200_weighted_quantile(ctx.h, pv, pw, n, mem, float(p), C.byref(out)))
    return out.value


def find_cutoff(target_pdf, xs, ws_or_q, q: Optional[float] = None, ctx: Optional[Context] = None) -> float:
    """find_cutoff(target_pdf, xs, q) / find_cutoff(target_pdf, xs, ws, q) -- confidence_set.jl:19-27; target_pdf is called once
    on the whole sample matrix (the batched closure) instead of column by column."""
    vals = target_pdf(xs)
    if q is None:
        return weighted_quantile(vals, None, 1.0 - float(ws_or_q), ctx)
    return weighted_quantile(vals, ws_or_q, 1.0 - q, ctx)


def approx_cutoff_area(target_pdf, xs, ws_or_c, c: Optional[float] = None, ctx: Optional[Context] = None) -> float:
    """approx_cutoff_area(target_pdf, xs, c) / (target_pdf, xs, ws, c) -- confidence_set.jl:47-56"""
    ctx = ctx or default_context()
    kv, pv, mem, n = _vec(target_pdf(xs))
    kw = pw = None
    if c is None:
        c = float(ws_or_c)
    else:
        kw, pw, m2, _ = _vec(ws_or_c)
        if m2 != mem:
            raise ValueError("values and weights must live in the same memory space")
    if mem == L.DEVICE:
        ctx.order_after_torch()
    out = C.c_double()
    ctx._check(ctx.lib.bosip_b200_cutoff_area(ctx.h, pv, pw, n, mem, float(c), C.byref(out)))
    return out.value


def set_iou(in_A, in_B, prior_logpdf_vals, ctx: Optional[Context] = None) -> float:
    """set_iou(in_A, in_B, x_prior, xs) -- confidence_set.jl:73-81 with logpdf(x_prior, xs) passed in (ws = 1 / pdf, normalised)."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(_host(in_A), dtype=np.uint8); b = np.ascontiguousarray(_host(in_B), dtype=np.uint8)
    lp = np.ascontiguousarray(_host(prior_logpdf_vals), dtype=np.float64)
    if not (a.size == b.size == lp.size):
        raise ValueError("in_A, in_B and the prior log-pdf values must have the same length")
    out = C.c_double()
    ctx._check(ctx.lib.bosip_b200_set_iou(ctx.h, a.ctypes.data, b.ctypes.data, lp.ctypes.data, a.size, L.HOST, C.byref(out)))
    return out.value


def confidence_iou(log_f_approx, log_f_expect, prior_logpdf_vals, q: float, ctx: Optional[Context] = None):
    """The core of calculate(::AEConfidence, bosip) (ae_confidence.jl:66-78) on the two closures' log-values: returns
    (iou, c_approx, c_expect).  One device call: two weighted quantiles + the IoU reduction."""
    ctx = ctx or default_context()
    ka, pa, mem, n = _vec(log_f_approx); kb, pb, m2, _ = _vec(log_f_expect)
    if _is_torch(ka) and not _is_torch(prior_logpdf_vals):
        import torch
        prior_logpdf_vals = torch.as_tensor(np.asarray(prior_logpdf_vals, dtype=np.float64), device=ka.device)
    kl, pl, m3, _ = _vec(prior_logpdf_vals)
    if not (mem == m2 == m3):
        raise ValueError("inputs must live in the same memory space")
    if mem == L.DEVICE:
        ctx.order_after_torch()
    iou, ca, cb = C.c_double(), C.c_double(), C.c_double()
    ctx._check(ctx.lib.bosip_b200_confidence_iou(ctx.h, pa, pb, pl, n, mem, float(q), C.byref(iou), C.byref(ca), C.byref(cb)))
    return iou.value, ca.value, cb.value


def _host(v):
    return v.detach().cpu().numpy() if _is_torch(v) else v


@dataclass
class AEConfidence:
    """AEConfidence(; max_iters, samples, xs, q, r) -- term_conds/ae_confidence.jl:19-47.  calculate() evaluates both closures on the
    whole sample matrix in ONE fused device sweep (the reference calls them column by column, :68-77) and reduces them on the device."""
    max_iters: Optional[int] = None
    samples: int = 10_000
    xs: Optional[np.ndarray] = None
    q: float = 0.95
    r: float = 0.95
    seed: int = 0

    def calculate(self, bosip) -> float:
        """calculate(cond, bosip) -- ae_confidence.jl:59-79: IoU of the q-confidence regions of the approximate and the expected posterior."""
        from .api import posterior_eval
        xs = self.xs if self.xs is not None else bosip.x_prior.rand(self.samples, np.random.default_rng(self.seed))
        post = bosip.model_posterior()
        res = posterior_eval(post, bosip.likelihood, bosip.x_prior, xs, L.WANT_APPROX | L.WANT_MEAN)
        xs_logpdf = bosip.x_prior.logpdf(_host(xs))
        ctx = (post[0] if isinstance(post, (list, tuple)) else post).ctx
        return confidence_iou(res["log_approx_post"], res["log_post_mean"], xs_logpdf, self.q, ctx)[0]

    def __call__(self, bosip) -> bool:
        """True = keep going (the IoU has not reached r yet), like every BosipTermCond (ae_confidence.jl:49-57)."""
        if self.max_iters is not None and getattr(bosip, "iterations", 0) >= self.max_iters:
            return False
        return self.calculate(bosip) < self.r


# ----------------------------------------------------------------------------------------------- OptMMD
@dataclass
class OptMMDMetric:
    """OptMMDMetric(; kernel, bounds, algorithm) -- metrics/opt_mmd.jl:17-31: the MMD with its kernel length-scales maximised
    inside [0, ub - lb] from x0 = (ub - lb) / 3 (:46-52); every objective evaluation is one device calc_mmd."""
    kernel: int
    bounds: tuple
    algorithm: str = "Nelder-Mead"
    options: dict = field(default_factory=dict)

    def calculate(self, true_samples, approx_samples, ctx: Optional[Context] = None) -> float:
        from scipy.optimize import minimize
        lb, ub = (np.asarray(b, dtype=np.float64) for b in self.bounds)
        span = ub - lb
        tiny = 1e-9 * span    # with_lengthscale needs lambda > 0; the reference's lower bound is 0 (exclusive in effect)

        def neg(lam):
            lam = np.clip(lam, tiny, span)
            return -calc_mmd(Kernel(self.kernel, lam), true_samples, approx_samples, ctx)

        sol = minimize(neg, span / 3.0, method=self.algorithm, bounds=list(zip(tiny, span)), options=self.options)
        self.lengthscales_ = np.clip(sol.x, tiny, span)
        return float(-neg(sol.x))
