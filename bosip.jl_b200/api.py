"""Host-side mirror of the BOSIP.jl interface for the batched evaluation path, on top of the C ABI.

Julia is not available in this image, so this module plays the role of the Julia glue (julia/BosipB200.jl): same names,
argument meaning, shapes and error behaviour as the reference (paths relative to /root/reference):

    NormalLikelihood                               src/likelihoods/normal.jl:16-19
    BosipProblem                                   src/types/problem.jl:28-35
    model_posterior -> B200GPPosterior             BOSS.model_posterior / BOSS.ModelPosterior (test/unit/utils.jl:3-11)
      mean / var / std / mean_and_var / mean_and_std (vector -> length p, matrix d x M -> p x M)
    log_approx_likelihood / log_likelihood_mean / log_likelihood_variance   src/posterior/likelihood.jl:19-59
    log_approx_posterior / log_posterior_mean / log_posterior_variance      src/posterior/log_posterior.jl:15-61
    approx_posterior / posterior_mean / posterior_variance (normalize=...)  src/posterior/posterior.jl:36-145
    evidence                                       src/posterior/evidence.jl:19-24
    MaxVar / LogMaxVar                             src/acquisitions/maxvar.jl, logmaxvar.jl
    B200CandidateAM.maximize_acquisition           BOSS SamplingAM / GridAM (docs/src/hyperparams.md:69)
    MMDMetric / TVMetric / calculate_metric / calc_mmd / calc_tv            src/metrics/mmd.jl, tv.jl

Every closure has the reference's two methods: a length-d vector -> scalar, a d x M matrix -> length-M vector.
Matrices use Julia indexing (X[k, i] = coordinate k of point i); NumPy inputs are converted to column-major once,
torch CUDA tensors are used in place (device pointers) when they are column-major (`X.t().is_contiguous()`)."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib as L

SE, MATERN32, MATERN52 = L.SE, L.MATERN32, L.MATERN52
UNIFORM, NORMAL, TRUNCNORMAL = L.PRIOR_UNIFORM, L.PRIOR_NORMAL, L.PRIOR_TRUNCNORMAL


class BosipB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[bosip_b200 {code}] {msg}")
        self.code = code


class DomainError(BosipB200Error, ValueError):
    """Julia's DomainError from `_assure_nonneg` (src/posterior/likelihood.jl:61-65)."""


class PosDefException(BosipB200Error):
    """Julia's PosDefException from `cholesky`."""


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(L.c_double_p)


def _colmajor(x, rows: Optional[int] = None):
    """Return (keepalive, address, mem, ncols) of a Julia-indexed (rows x n) matrix or length-n vector."""
    if _is_torch(x):
        import torch
        if x.dtype != torch.float64:
            raise TypeError("device tensors must be float64")
        if x.dim() == 2:
            if not x.t().is_contiguous():
                x = x.t().contiguous().t()
            n = x.shape[1]
        else:
            x = x.contiguous(); n = x.shape[0]
        if rows is not None and x.dim() == 2 and x.shape[0] != rows:
            raise ValueError(f"expected {rows} rows, got {x.shape[0]}")
        mem = L.DEVICE if x.is_cuda else L.HOST
        return x, x.data_ptr(), mem, n
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 2:
        if rows is not None and a.shape[0] != rows:
            raise ValueError(f"expected {rows} rows, got {a.shape[0]}")
        a = np.asfortranarray(a); n = a.shape[1]
    else:
        a = np.ascontiguousarray(a); n = a.shape[0]
    return a, a.ctypes.data, L.HOST, n


def _out(shape, mem, like=None):
    """Allocate an output in the same memory space as the inputs; 2-D outputs are column-major."""
    if mem == L.DEVICE:
        import torch
        if len(shape) == 2:
            t = torch.empty((shape[1], shape[0]), dtype=torch.float64, device=like.device).t()
        else:
            t = torch.empty(shape, dtype=torch.float64, device=like.device)
        return t, t.data_ptr()
    a = np.empty(shape, dtype=np.float64, order="F")
    return a, a.ctypes.data


class Context:
    """One per (process, GPU).  Raises when the library or a B200 is missing -- there is no CPU path."""

    def __init__(self, device: int = 0):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.bosip_b200_init(device, C.byref(h))
        if rc != 0:
            raise BosipB200Error(rc, (self.lib.bosip_b200_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def _check(self, rc: int):
        if rc == 0:
            return
        msg = (self.lib.bosip_b200_last_error(self.h) or b"").decode()
        if rc == L.EDOMAIN:
            raise DomainError(rc, msg)
        if rc == L.ENOTPD:
            raise PosDefException(rc, msg)
        raise BosipB200Error(rc, msg)

    def set_stream(self, cuda_stream: Optional[int]):
        """Launch compute kernels on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream); None restores."""
        self._check(self.lib.bosip_b200_set_stream(self.h, cuda_stream))

    def order_after_torch(self):
        """Order the compute stream after torch's CURRENT stream on this device.  Every call that hands device pointers to the
        library goes through this first: the tensors may still be being written (a `.t().contiguous()` copy, `torch.zeros`,
        `copy_()`), or their memory may have been recycled by torch's caching allocator while earlier torch kernels that use it
        are still queued -- the library's own streams are non-blocking and see none of that implicitly."""
        import torch
        self._check(self.lib.bosip_b200_stream_wait(self.h, torch.cuda.current_stream(self.device).cuda_stream))

    def i8_issuers(self) -> int:
        """MMA-issuing threads of the INT8 engine: 2 after the bitwise self-check passed, 1 after a mismatch, 0 before the first INT8 launch."""
        n = C.c_int()
        self._check(self.lib.bosip_b200_i8_issuers(self.h, C.byref(n)))
        return n.value

    def variance_engine(self) -> dict:
        """Current variance-engine setting: {'engine': 'fp64'|'int8'|'auto', 'slices': bytes of alpha^2 L^-1, 'kstar_slices': bytes of K*,
        'slice_pairs': INT8 GEMMs per contraction}."""
        e, s, a = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.bosip_b200_get_variance_engine(self.h, C.byref(e), C.byref(s), C.byref(a)))
        S, SA = s.value, a.value
        return {"engine": {0: "fp64", 1: "int8", 2: "auto"}[e.value], "slices": S, "kstar_slices": SA, "slice_pairs": sum(S - i for i in range(SA))}

    def i8_pair(self) -> int:
        """Paired (cta_group::2) INT8 kernel: 1 in use (BOSIP_B200_I8_PAIR=1 and the bitwise self-check passed), -1 self-check failed, 0 otherwise."""
        n = C.c_int()
        self._check(self.lib.bosip_b200_i8_pair(self.h, C.byref(n)))
        return n.value

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self.lib.bosip_b200_launch_count(self.h, C.byref(n)))
        return n.value

    def set_chunk(self, max_points: int):
        self._check(self.lib.bosip_b200_set_chunk(self.h, int(max_points)))

    def set_variance_engine(self, engine: str = "auto", slices: int = 0):
        """'fp64' (DMMA), 'int8' (tcgen05 INT8 on error-free byte slices; `slices` in 6..8, 0 = 7) or 'auto' (int8 for N >= 128)."""
        code = {"fp64": L.VAR_FP64, "int8": L.VAR_INT8, "auto": L.VAR_AUTO}[engine]
        self._check(self.lib.bosip_b200_set_variance_engine(self.h, code, int(slices)))

    def int8_peak(self, sustained: bool = False):
        """Measured tcgen05 INT8 issue-rate ceiling (TOPS, 2 ops per MAC) of this GPU: the INT8 engine's roofline denominator.
        Returns the burst figure, or (burst, sustained-under-power-cap) when `sustained` (takes ~2 s)."""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.bosip_b200_int8_peak(self.h, C.byref(a), C.byref(b) if sustained else None))
        return (a.value, b.value) if sustained else a.value

    def int8_peak_random(self) -> float:
        """Sustained (~2 s, power-capped) tcgen05 INT8 issue rate with pseudo-random operand bytes (TOPS): what the multipliers' operand-dependent
        switching energy leaves of the issue-rate ceiling in a long run."""
        a = C.c_double()
        self._check(self.lib.bosip_b200_int8_peak_random(self.h, C.byref(a)))
        return a.value

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.bosip_b200_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return {"dmma_tflops": a.value, "dfma_tflops": b.value}

    def timing_reset(self, enable: bool = True):
        self._check(self.lib.bosip_b200_timing_reset(self.h, 1 if enable else 0))

    def timing_get(self):
        ms = (C.c_double * 3)(); n = (C.c_int64 * 3)()
        self._check(self.lib.bosip_b200_timing_get(self.h, ms, n))
        names = ("kstar_mean", "var_gemm", "epilogue")
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(names)}

    def close(self):
        if getattr(self, "h", None):
            self.lib.bosip_b200_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


# ----------------------------------------------------------------------------------------------- problem pieces
@dataclass
class NormalLikelihood:
    """NormalLikelihood(; z_obs, std_obs) -- src/likelihoods/normal.jl:16-19"""
    z_obs: np.ndarray
    std_obs: np.ndarray

    def __post_init__(self):
        self.z_obs = np.ascontiguousarray(self.z_obs, dtype=np.float64)
        self.std_obs = np.ascontiguousarray(self.std_obs, dtype=np.float64)
        if self.z_obs.shape != self.std_obs.shape:
            raise ValueError("z_obs and std_obs must have the same length")

    @property
    def obs_dim(self): return len(self.z_obs)
    @property
    def model_dim(self): return len(self.z_obs)

    def _c(self):
        return L.LikeNormal(len(self.z_obs), L.LIKE_NORMAL, _dptr(self.z_obs), _dptr(self.std_obs), None)


GaussianLikelihood = NormalLikelihood  # src/likelihoods/normal.jl:26


@dataclass
class LogNormalLikelihood:
    """LogNormalLikelihood(; log_z_obs, CV) -- src/likelihoods/lognormal.jl:26-42; the model approximates log(y)."""
    log_z_obs: np.ndarray
    CV: np.ndarray

    def __post_init__(self):
        self.log_z_obs = np.ascontiguousarray(self.log_z_obs, dtype=np.float64)
        self.CV = np.ascontiguousarray(self.CV, dtype=np.float64)
        assert self.log_z_obs.shape == self.CV.shape

    @property
    def obs_dim(self): return len(self.log_z_obs)
    @property
    def model_dim(self): return len(self.log_z_obs)

    def _c(self):
        return L.LikeNormal(len(self.log_z_obs), L.LIKE_LOGNORMAL, _dptr(self.log_z_obs), _dptr(self.CV), None)


@dataclass
class NormalDiffLikelihood:
    """NormalDiffLikelihood(; std_obs) -- src/likelihoods/normal_diff.jl:17-19; the model approximates |y - z_obs|."""
    std_obs: np.ndarray

    def __post_init__(self):
        self.std_obs = np.ascontiguousarray(self.std_obs, dtype=np.float64)

    @property
    def obs_dim(self): return len(self.std_obs)
    @property
    def model_dim(self): return len(self.std_obs)

    def _c(self):
        return L.LikeNormal(len(self.std_obs), L.LIKE_NORMAL_DIFF, None, _dptr(self.std_obs), None)


@dataclass
class NormalSumLikelihood:
    """NormalSumLikelihood(; sum_lengths, z_obs, std_obs) -- src/likelihoods/normal_sum.jl:18-27: every observation is the sum
    of a contiguous group of model outputs."""
    sum_lengths: np.ndarray
    z_obs: np.ndarray
    std_obs: np.ndarray

    def __post_init__(self):
        self.sum_lengths = np.ascontiguousarray(self.sum_lengths, dtype=np.int32)
        self.z_obs = np.ascontiguousarray(self.z_obs, dtype=np.float64)
        self.std_obs = np.ascontiguousarray(self.std_obs, dtype=np.float64)
        assert len(self.sum_lengths) == len(self.z_obs), "sum_lengths and z_obs must have the same length"

    @property
    def obs_dim(self): return len(self.z_obs)
    @property
    def model_dim(self): return int(self.sum_lengths.sum())

    def _c(self):
        return L.LikeNormal(len(self.z_obs), L.LIKE_NORMAL_SUM, _dptr(self.z_obs), _dptr(self.std_obs),
                            self.sum_lengths.ctypes.data_as(L.c_int32_p))


@dataclass
class ExpLikelihood:
    """ExpLikelihood() -- src/likelihoods/exp.jl:7: the (single) model output is the log-likelihood itself."""

    @property
    def obs_dim(self): return 1
    @property
    def model_dim(self): return 1

    def _c(self):
        return L.LikeNormal(1, L.LIKE_EXP, None, None, None)


@dataclass
class ProductPrior:
    """Product of univariate priors: kind[k] in {UNIFORM (a,b), NORMAL (m,s), TRUNCNORMAL (m,s,a,b)}.
    Stands for the `x_prior::MultivariateDistribution` of BosipProblem (src/types/problem.jl:33) for the priors the
    reference's examples use; any other prior is passed as precomputed log-pdf values (`logprior=`)."""
    kind: Sequence[int]
    params: Sequence[Sequence[float]]

    def __post_init__(self):
        self.kind = np.ascontiguousarray(self.kind, dtype=np.int32)
        prm = np.zeros((len(self.kind), 4))
        for k, q in enumerate(self.params):
            prm[k, :len(q)] = q
        self.params = np.ascontiguousarray(prm)

    @property
    def d(self): return len(self.kind)

    def _c(self, logprior_ptr=None):
        return L.Prior(self.d, self.kind.ctypes.data_as(L.c_int32_p), _dptr(self.params), logprior_ptr)

    def logpdf(self, X) -> np.ndarray:
        """logpdf(x_prior, x) per column of X (d x M), on the host: -Inf outside the support (Distributions semantics;
        the device epilogue applies the same formulas inside the fused sweep, src/posterior/log_posterior.jl:228-229)."""
        from scipy.special import erfc
        X = np.asarray(X, dtype=np.float64).reshape(self.d, -1)
        out = np.zeros(X.shape[1])
        for k in range(self.d):
            q, x = self.params[k], X[k]
            if self.kind[k] == UNIFORM:
                out += np.where((x >= q[0]) & (x <= q[1]), -math.log(q[1] - q[0]), -np.inf)
            else:
                lp = -0.5 * ((x - q[0]) / q[1]) ** 2 - 0.5 * math.log(2.0 * math.pi) - math.log(q[1])
                if self.kind[k] == TRUNCNORMAL:
                    tp = 0.5 * erfc(-(q[3] - q[0]) / q[1] / math.sqrt(2.0)) - 0.5 * erfc(-(q[2] - q[0]) / q[1] / math.sqrt(2.0))
                    lp = np.where((x >= q[2]) & (x <= q[3]), lp - math.log(tp), -np.inf)
                out += lp
        return out

    def rand(self, n: int, rng: np.random.Generator) -> np.ndarray:
        """d x n samples (rand(x_prior, n), src/posterior/evidence.jl:21)."""
        from scipy.stats import truncnorm
        out = np.empty((self.d, n))
        for k in range(self.d):
            q = self.params[k]
            if self.kind[k] == UNIFORM:
                out[k] = rng.uniform(q[0], q[1], n)
            elif self.kind[k] == NORMAL:
                out[k] = rng.normal(q[0], q[1], n)
            else:
                out[k] = truncnorm.rvs((q[2] - q[0]) / q[1], (q[3] - q[0]) / q[1], loc=q[0], scale=q[1], size=n, random_state=rng)
        return out


@dataclass
class GaussianProcess:
    """BOSS.GaussianProcess with fitted (MAP) hyper-parameters: kernel, lambda d x p, alpha p, sigma p
    (src/subset_problem.jl:38-44).  `mean` is an optional callable X (d x M) -> p x M (prior mean function)."""
    kernel: int
    lengthscales: np.ndarray
    amplitudes: np.ndarray
    noise_std: np.ndarray
    mean: Optional[Callable] = None
    jitter: float = 0.0
    pred_noise: bool = False
    clamp_var: bool = True


@dataclass
class BosipProblem:
    """The fields of BosipProblem the evaluation path reads (src/types/problem.jl:28-35): data, model, likelihood, prior."""
    X: np.ndarray            # d x N
    Y: np.ndarray            # p x N
    model: GaussianProcess
    likelihood: NormalLikelihood
    x_prior: ProductPrior
    bounds: Optional[tuple] = None   # (lb, ub) of the domain
    ctx: Optional[Context] = None
    _post: Optional["B200GPPosterior"] = field(default=None, repr=False)

    def model_posterior(self):
        """BOSS.model_posterior(problem): one posterior for MAP parameters (UniFittedParams), a list of posteriors when
        `model` is a list of hyper-parameter samples (MultiFittedParams, src/posterior/log_posterior.jl:86-89)."""
        if self._post is None:
            if isinstance(self.model, (list, tuple)):
                self._post = [model_posterior(m, self.X, self.Y, ctx=self.ctx) for m in self.model]
            else:
                self._post = model_posterior(self.model, self.X, self.Y, ctx=self.ctx)
        return self._post

    def augment(self, x_new: np.ndarray, y_new: np.ndarray):
        """BOSS.augment_dataset!: append simulations; the cached posterior is dropped."""
        self.X = np.concatenate([self.X, np.asarray(x_new, float).reshape(self.X.shape[0], -1)], axis=1)
        self.Y = np.concatenate([self.Y, np.asarray(y_new, float).reshape(self.Y.shape[0], -1)], axis=1)
        for q in (self._post if isinstance(self._post, list) else [self._post]):
            if q is not None:
                q.free()
        self._post = None


# ----------------------------------------------------------------------------------------------- ModelPosterior seam
class B200GPPosterior:
    """BOSS.ModelPosterior backed by a device handle.  mean/var/... follow test/unit/utils.jl:8-11: vector -> p, matrix -> p x M."""

    def __init__(self, ctx: Context, handle, model: GaussianProcess, d: int, N: int, p: int):
        self.ctx, self.h, self.model, self.d, self.N, self.p = ctx, handle, model, d, N, p

    def free(self):
        if self.h:
            self.ctx.lib.bosip_b200_gp_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _mean_at(self, X):
        if self.model.mean is None:
            return None, None
        m = np.asfortranarray(np.asarray(self.model.mean(np.asarray(X)), dtype=np.float64).reshape(self.p, -1))
        return m, m.ctypes.data

    def _predict(self, X, want_mu=True, want_var=True):
        vec = (np.ndim(X) == 1) if not _is_torch(X) else (X.dim() == 1)
        Xm = X.reshape(self.d, 1) if vec else X
        keep, addr, mem, M = _colmajor(Xm, self.d)
        mkeep, maddr = self._mean_at(Xm) if mem == L.HOST else (None, None)
        if self.model.mean is not None and mem == L.DEVICE:
            raise NotImplementedError("a prior mean function needs host inputs")
        mu, mu_p = _out((self.p, M), mem, keep) if want_mu else (None, None)
        var, var_p = _out((self.p, M), mem, keep) if want_var else (None, None)
        if mem == L.DEVICE:
            self.ctx.order_after_torch()
        self.ctx._check(self.ctx.lib.bosip_b200_gp_predict(self.h, addr, M, mem, maddr, mu_p, var_p))
        if vec:
            mu = mu[:, 0] if mu is not None else None
            var = var[:, 0] if var is not None else None
        return mu, var

    def mean(self, X): return self._predict(X, True, False)[0]
    def var(self, X): return self._predict(X, False, True)[1]
    def std(self, X): return np.sqrt(self.var(X)) if not _is_torch(X) else self.var(X).sqrt()
    def mean_and_var(self, X): return self._predict(X, True, True)

    def mean_and_std(self, X):
        mu, var = self._predict(X, True, True)
        return mu, (np.sqrt(var) if not _is_torch(var) else var.sqrt())

    def factors(self):
        """(L p x N x N, Linv p x N x N, a N x p) as the device holds them."""
        Lm = np.empty((self.p, self.N, self.N)); Li = np.empty_like(Lm); a = np.empty((self.p, self.N))
        self.ctx._check(self.ctx.lib.bosip_b200_gp_get_factors(self.h, _dptr(Lm), _dptr(Li), _dptr(a)))
        # device layout is column-major per output: transpose the trailing axes to get L[j][r, c]
        return Lm.transpose(0, 2, 1).copy(), Li.transpose(0, 2, 1).copy(), a.T.copy()


def _opts(model: GaussianProcess):
    return L.GpOpts(float(model.jitter), int(model.pred_noise), int(model.clamp_var))


def model_posterior(model: GaussianProcess, X, Y, ctx: Optional[Context] = None) -> B200GPPosterior:
    """BOSS.model_posterior(problem) -- conditions the GP on the device (src/posterior/log_posterior.jl:81)."""
    ctx = ctx or default_context()
    X = np.asfortranarray(np.asarray(X, dtype=np.float64)); Y = np.asfortranarray(np.asarray(Y, dtype=np.float64))
    d, N = X.shape; p = Y.shape[0]
    lam = np.asfortranarray(np.asarray(model.lengthscales, dtype=np.float64).reshape(d, p))
    alpha = np.ascontiguousarray(model.amplitudes, dtype=np.float64); sigma = np.ascontiguousarray(model.noise_std, dtype=np.float64)
    mX = None
    if model.mean is not None:
        mX = np.asfortranarray(np.asarray(model.mean(X), dtype=np.float64).reshape(p, N))
    h = C.c_void_p(); o = _opts(model)
    ctx._check(ctx.lib.bosip_b200_gp_fit(ctx.h, model.kernel, d, N, p, _dptr(X), _dptr(Y), _dptr(lam), _dptr(alpha), _dptr(sigma),
                                         _dptr(mX) if mX is not None else None, C.byref(o), C.byref(h)))
    return B200GPPosterior(ctx, h, model, d, N, p)


def model_posterior_from_factors(model: GaussianProcess, X, L_fac, a, ctx: Optional[Context] = None) -> B200GPPosterior:
    """Host-factorised variant (exact-parity runs): L_fac[j] lower Cholesky factor of K_j, a N x p."""
    ctx = ctx or default_context()
    X = np.asfortranarray(np.asarray(X, dtype=np.float64))
    d, N = X.shape
    L_fac = np.asarray(L_fac, dtype=np.float64); p = L_fac.shape[0]
    Lc = np.ascontiguousarray(L_fac.transpose(0, 2, 1))          # per-output column-major
    ac = np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)  # N x p column-major == [p][N]
    lam = np.asfortranarray(np.asarray(model.lengthscales, dtype=np.float64).reshape(d, p))
    alpha = np.ascontiguousarray(model.amplitudes, dtype=np.float64); sigma = np.ascontiguousarray(model.noise_std, dtype=np.float64)
    h = C.c_void_p(); o = _opts(model)
    ctx._check(ctx.lib.bosip_b200_gp_load_factors(ctx.h, model.kernel, d, N, p, _dptr(X), _dptr(lam), _dptr(alpha), _dptr(sigma),
                                                  _dptr(Lc), _dptr(ac), C.byref(o), C.byref(h)))
    return B200GPPosterior(ctx, h, model, d, N, p)


# ----------------------------------------------------------------------------------------------- fused closures
def posterior_eval(post: B200GPPosterior, like: NormalLikelihood, prior: Optional[ProductPrior], X, want: int,
                   logprior=None, acq=None, index_offset: int = 0, out=None):
    """One fused call: returns dict of the requested length-M outputs + 'n_clamped'; with `acq` (MaxVar()/LogMaxVar()) also
    'best_idx' (index_offset + position) and 'best_val' from the same sweep.  `out` may hold preallocated output buffers
    (dict name -> array/tensor in the same memory space as X)."""
    ensemble = list(post) if isinstance(post, (list, tuple)) else None   # MultiFittedParams: hyper-parameter samples
    if ensemble is not None:
        post = ensemble[0]
    vec = (np.ndim(X) == 1) if not _is_torch(X) else (X.dim() == 1)
    Xm = X.reshape(post.d, 1) if vec else X
    keep, addr, mem, M = _colmajor(Xm, post.d)
    models = [q.model for q in (ensemble or [post])]
    if any(m.mean is not None for m in models):
        # the C ABI takes ONE p x M array of prior-mean values, evaluated on the host
        if mem == L.DEVICE:
            raise NotImplementedError("a prior mean function needs host inputs (the mean is a host callable)")
        if any(m.mean is not models[0].mean for m in models):
            raise NotImplementedError("hyper-parameter samples with different prior mean functions are not supported")
    mkeep, maddr = post._mean_at(Xm) if mem == L.HOST else (None, None)
    lk = like._c()
    lp_keep = lp_addr = None
    if logprior is not None:
        lp_keep, lp_addr, lp_mem, _ = _colmajor(logprior)
        if lp_mem != mem:
            raise ValueError("logprior must live in the same memory space as X")
    if prior is not None:
        pr = prior._c(lp_addr)
    else:
        pr = L.Prior(post.d, None, None, lp_addr)
    outs = {}
    ptrs = []
    for bit, name in ((L.WANT_APPROX, "log_approx_post"), (L.WANT_MEAN, "log_post_mean"), (L.WANT_VAR, "log_post_var")):
        if want & bit:
            if out is not None and name in out:
                o = out[name]; p_ = o.data_ptr() if _is_torch(o) else o.ctypes.data
            else:
                o, p_ = _out((M,), mem, keep)
            outs[name] = o; ptrs.append(p_)
        else:
            ptrs.append(None)
    ncl = C.c_int64(0); bi = C.c_int64(-1); bv = C.c_double(0.0)
    if mem == L.DEVICE:
        post.ctx.order_after_torch()
    if ensemble is None:
        post.ctx._check(post.ctx.lib.bosip_b200_posterior_eval_argmax(
            post.h, C.byref(lk), C.byref(pr), addr, M, mem, maddr, want, ptrs[0], ptrs[1], ptrs[2], C.byref(ncl),
            acq.acq_id if acq is not None else -1, index_offset, C.byref(bi), C.byref(bv)))
    else:
        hs = (C.c_void_p * len(ensemble))(*[q.h for q in ensemble])
        post.ctx._check(post.ctx.lib.bosip_b200_posterior_eval_bi(
            hs, len(ensemble), C.byref(lk), C.byref(pr), addr, M, mem, maddr, want, ptrs[0], ptrs[1], ptrs[2], C.byref(ncl),
            acq.acq_id if acq is not None else -1, index_offset, C.byref(bi), C.byref(bv)))
    if vec:
        outs = {k: float(v[0]) for k, v in outs.items()}
    outs["n_clamped"] = ncl.value
    if acq is not None:
        outs["best_idx"], outs["best_val"] = bi.value, bv.value
    return outs


def _closure(post, like, prior, want, key, transform=None):
    def f(X, logprior=None):
        v = posterior_eval(post, like, prior, X, want, logprior)[key]
        return transform(v) if transform else v
    return f


# ----------------------------------------------------------------------------------------------- Likelihood API layer
_LIKE_OUTS = ("loglike_marginal", "loglike", "log_marginal_mean", "log_like_mean", "log_sq_like_mean", "log_like_var")


def normal_likelihood_eval(like: NormalLikelihood, mu, var=None, outputs=_LIKE_OUTS, ctx: Optional[Context] = None):
    """NormalLikelihood algebra on the device for ANY ModelPosterior's (mu, var): p (vector) or p x M (matrix) inputs.
    Returns a dict with the requested outputs (marginal ones p / p x M, the others scalar / length M) + 'n_clamped'."""
    ctx = ctx or default_context()
    p, nq = like.model_dim, like.obs_dim
    vec = (np.ndim(mu) == 1) if not _is_torch(mu) else (mu.dim() == 1)
    mu_m = mu.reshape(p, 1) if vec else mu
    kmu, pmu, mem, M = _colmajor(mu_m, p)
    kvar = pvar = None
    if var is not None:
        var_m = var.reshape(p, 1) if vec else var
        kvar, pvar, mem2, _ = _colmajor(var_m, p)
        if mem2 != mem:
            raise ValueError("mu and var must live in the same memory space")
    res, ptrs = {}, []
    for name in _LIKE_OUTS:
        if name in outputs:
            shape = (nq, M) if name in ("loglike_marginal", "log_marginal_mean") else (M,)
            o, addr = _out(shape, mem, kmu); res[name] = o; ptrs.append(addr)
        else:
            ptrs.append(None)
    ncl = C.c_int64(0); lk = like._c()
    if mem == L.DEVICE:
        ctx.order_after_torch()
    ctx._check(ctx.lib.bosip_b200_normal_likelihood_eval(ctx.h, C.byref(lk), pmu, pvar, M, mem, *ptrs, C.byref(ncl)))
    if vec:
        res = {k: (v[:, 0] if v.ndim == 2 else float(v[0])) for k, v in res.items()}
    res["n_clamped"] = ncl.value
    return res


def loglike_marginal(like: NormalLikelihood, Y, x=None):
    """src/likelihoods/normal.jl:28-30 + matrix fallback src/types/likelihood.jl:53-54"""
    return normal_likelihood_eval(like, Y, None, ("loglike_marginal",))["loglike_marginal"]


def loglike(like: NormalLikelihood, Y, x=None):
    """src/types/likelihood.jl:66-71"""
    return normal_likelihood_eval(like, Y, None, ("loglike",))["loglike"]


def _generic(like, post, out_name, need_var):
    """Closure over an arbitrary ModelPosterior (anything with mean(X) / var(X), e.g. the reference's MockModelPosterior)."""
    def f(X):
        mu = post.mean(X)
        var = post.var(X) if need_var else None
        return normal_likelihood_eval(like, mu, var, (out_name,))[out_name]
    return f


def log_approx_marginal_likelihood(like: NormalLikelihood, post):
    """src/posterior/likelihood.jl:6-16"""
    return _generic(like, post, "loglike_marginal", False)


def log_marginal_likelihood_mean(like: NormalLikelihood, post):
    """src/likelihoods/normal.jl:32-53"""
    return _generic(like, post, "log_marginal_mean", True)


def log_sq_likelihood_mean(like: NormalLikelihood, post):
    """src/likelihoods/normal.jl:55-79"""
    return _generic(like, post, "log_sq_like_mean", True)


# likelihood-level closures (no prior): src/posterior/likelihood.jl:19-59.  A device-backed GP posterior takes the fused
# path (one ccall per X); any other ModelPosterior goes through its own mean/var and the Likelihood-API kernel.
def _device_backed(post) -> bool:
    return isinstance(post, B200GPPosterior) or (isinstance(post, (list, tuple)) and all(isinstance(q, B200GPPosterior) for q in post))


def log_approx_likelihood(like: NormalLikelihood, post):
    if _device_backed(post):
        return _closure(post, like, None, L.WANT_APPROX, "log_approx_post")
    return _generic(like, post, "loglike", False)


def log_likelihood_mean(like: NormalLikelihood, post):
    if _device_backed(post):
        return _closure(post, like, None, L.WANT_MEAN, "log_post_mean")
    return _generic(like, post, "log_like_mean", True)


def log_likelihood_variance(like: NormalLikelihood, post):
    if _device_backed(post):
        return _closure(post, like, None, L.WANT_VAR, "log_post_var")
    return _generic(like, post, "log_like_var", True)


# posterior-level closures: src/posterior/log_posterior.jl:15-61
def log_approx_posterior(bosip: BosipProblem):
    return _closure(bosip.model_posterior(), bosip.likelihood, bosip.x_prior, L.WANT_APPROX, "log_approx_post")


def log_posterior_mean(bosip: BosipProblem):
    return _closure(bosip.model_posterior(), bosip.likelihood, bosip.x_prior, L.WANT_MEAN, "log_post_mean")


def log_posterior_variance(bosip: BosipProblem):
    return _closure(bosip.model_posterior(), bosip.likelihood, bosip.x_prior, L.WANT_VAR, "log_post_var")


def _exp(v):
    return v.exp() if _is_torch(v) else (math.exp(v) if np.isscalar(v) else np.exp(v))


def evidence(bosip: BosipProblem, which: str = "approx", xs=None, samples: int = 10_000, rng=None) -> float:
    """evidence(post, x_prior; xs, samples) -- src/posterior/evidence.jl:19-24: mean_i post(x_i)/pdf(prior, x_i)."""
    post = bosip.model_posterior()
    ensemble = list(post) if isinstance(post, (list, tuple)) else [post]     # MultiFittedParams: averaged in linear space
    post = ensemble[0]
    if any(q.model.mean is not None for q in ensemble):
        raise NotImplementedError("evidence with a prior mean function is not wired through the fused sweep")
    if xs is None:
        xs = bosip.x_prior.rand(samples, rng or np.random.default_rng())
    keep, addr, mem, S = _colmajor(xs, post.d)
    lk = bosip.likelihood._c(); pr = bosip.x_prior._c()
    out = C.c_double()
    w = L.WANT_APPROX if which == "approx" else L.WANT_MEAN
    if mem == L.DEVICE:
        post.ctx.order_after_torch()
    hs = (C.c_void_p * len(ensemble))(*[q.h for q in ensemble])
    post.ctx._check(post.ctx.lib.bosip_b200_evidence_sum_bi(hs, len(ensemble), C.byref(lk), C.byref(pr), addr, S, mem, w, C.byref(out)))
    return out.value / S


def _normalised(bosip, log_f, which, power, normalize, xs, samples):
    # src/posterior/posterior.jl:36-50, 84-98, 130-145
    if not normalize:
        return lambda X, **kw: _exp(log_f(X, **kw))
    log_py = math.log(evidence(bosip, which, xs, samples))
    return lambda X, **kw: _exp(log_f(X, **kw) - power * log_py)


def approx_posterior(bosip: BosipProblem, normalize=False, xs=None, samples=10_000):
    return _normalised(bosip, log_approx_posterior(bosip), "approx", 1.0, normalize, xs, samples)


def posterior_mean(bosip: BosipProblem, normalize=False, xs=None, samples=10_000):
    return _normalised(bosip, log_posterior_mean(bosip), "mean", 1.0, normalize, xs, samples)


def posterior_variance(bosip: BosipProblem, normalize=False, xs=None, samples=10_000):
    # normalised variance divides by p(z)^2 with p(z) estimated from the posterior MEAN (posterior.jl:135-138)
    return _normalised(bosip, log_posterior_variance(bosip), "mean", 2.0, normalize, xs, samples)


def approx_likelihood(bosip: BosipProblem):
    f = log_approx_likelihood(bosip.likelihood, bosip.model_posterior()); return lambda X: _exp(f(X))


def likelihood_mean(bosip: BosipProblem):
    f = log_likelihood_mean(bosip.likelihood, bosip.model_posterior()); return lambda X: _exp(f(X))


def likelihood_variance(bosip: BosipProblem):
    f = log_likelihood_variance(bosip.likelihood, bosip.model_posterior()); return lambda X: _exp(f(X))


# ----------------------------------------------------------------------------------------------- acquisitions
class MaxVar:
    """src/acquisitions/maxvar.jl:7-11: acq(bosip) = posterior_variance(bosip; normalize=false)"""
    acq_id = L.ACQ_MAXVAR

    def __call__(self, bosip: BosipProblem, options=None):
        return posterior_variance(bosip, normalize=False)


class LogMaxVar:
    """src/acquisitions/logmaxvar.jl:11-15: acq(bosip) = log_posterior_variance(bosip)"""
    acq_id = L.ACQ_LOGMAXVAR

    def __call__(self, bosip: BosipProblem, options=None):
        return log_posterior_variance(bosip)


@dataclass
class CandidateSource:
    """points (d x M, host or device) | philox(seed, lb, ub, count) | grid(lb, ub, n_per_dim)"""
    kind: int
    points: object = None
    seed: int = 0
    lb: Optional[np.ndarray] = None
    ub: Optional[np.ndarray] = None
    n_per_dim: int = 0
    count: int = 0

    @staticmethod
    def from_points(X): return CandidateSource(L.CAND_POINTS, points=X, count=(X.shape[1]))
    @staticmethod
    def philox(seed, lb, ub, count): return CandidateSource(L.CAND_PHILOX, seed=int(seed), lb=np.ascontiguousarray(lb, float), ub=np.ascontiguousarray(ub, float), count=int(count))
    @staticmethod
    def grid(lb, ub, n_per_dim):
        lb = np.ascontiguousarray(lb, float)
        return CandidateSource(L.CAND_GRID, lb=lb, ub=np.ascontiguousarray(ub, float), n_per_dim=int(n_per_dim), count=int(n_per_dim) ** len(lb))


def acq_argmax(post: B200GPPosterior, like, prior, acq, src: CandidateSource, start: int = 0, count: Optional[int] = None):
    """Fused evaluate + argmax over candidates start..start+count-1 of `src`.  Returns (global index, value, x)."""
    ensemble = list(post) if isinstance(post, (list, tuple)) else [post]     # MultiFittedParams
    post = ensemble[0]
    if any(q.model.mean is not None for q in ensemble):
        raise NotImplementedError("acq_argmax with a prior mean function: evaluate posterior_eval(..., acq=...) on explicit host points")
    count = src.count - start if count is None else count
    cs = L.CandSrc(); cs.kind = src.kind; keep = None
    if src.kind == L.CAND_POINTS:
        pts = src.points[:, start:start + count]
        keep, addr, mem, _ = _colmajor(pts, post.d)
        cs.points = addr; cs.mem = mem
        if mem == L.DEVICE:
            post.ctx.order_after_torch()
    else:
        cs.seed = src.seed; cs.lb = _dptr(src.lb); cs.ub = _dptr(src.ub); cs.n_per_dim = src.n_per_dim
    lk = like._c(); pr = prior._c() if prior is not None else L.Prior(post.d, None, None, None)
    bi, bv = C.c_int64(-1), C.c_double()
    bx = np.full(post.d, np.nan)
    hs = (C.c_void_p * len(ensemble))(*[q.h for q in ensemble])
    post.ctx._check(post.ctx.lib.bosip_b200_acq_argmax_bi(hs, len(ensemble), C.byref(lk), C.byref(pr), C.byref(cs), count, start, acq.acq_id,
                                                         C.byref(bi), C.byref(bv), _dptr(bx)))
    return bi.value, bv.value, bx


class B200CandidateAM:
    """BOSS.AcquisitionMaximizer over a candidate set (SamplingAM / GridAM semantics): x* = cand[argmax acq.(cand)],
    first maximal index on ties.  With torch.distributed initialised the candidate range is sharded across ranks and
    the per-rank winners are all-gathered (see parallel.py)."""

    def __init__(self, source: CandidateSource):
        self.source = source

    def maximize_acquisition(self, bosip: BosipProblem, acq):
        from . import parallel
        post = bosip.model_posterior()
        start, count = parallel.shard_range(self.source.count)
        idx, val, x = acq_argmax(post, bosip.likelihood, bosip.x_prior, acq, self.source, start, count)
        idx, val, x = parallel.argmax_allgather(idx, val, x)
        if idx < 0:
            raise ValueError("maximize_acquisition: the candidate set is empty")
        return x, val


def candidates(ctx: Context, src: CandidateSource, start: int, count: int, device_out=None) -> np.ndarray:
    """Materialise PHILOX/GRID candidates (d x count)."""
    d = len(src.lb)
    cs = L.CandSrc(); cs.kind = src.kind; cs.seed = src.seed; cs.lb = _dptr(src.lb); cs.ub = _dptr(src.ub); cs.n_per_dim = src.n_per_dim
    if device_out is not None:
        ctx.order_after_torch()
        ctx._check(ctx.lib.bosip_b200_candidates(ctx.h, C.byref(cs), d, count, start, L.DEVICE, device_out.data_ptr()))
        return device_out
    out = np.empty((d, count), order="F")
    ctx._check(ctx.lib.bosip_b200_candidates(ctx.h, C.byref(cs), d, count, start, L.HOST, out.ctypes.data))
    return out


# ----------------------------------------------------------------------------------------------- metrics
@dataclass
class Kernel:
    """KernelFunctions kernel with ARD length-scales: with_lengthscale(kernel, lengthscales) (src/metrics/opt_mmd.jl:39)."""
    kernel: int
    lengthscales: np.ndarray


def calc_mmd(kernel: Kernel, X, Y, ctx: Optional[Context] = None) -> float:
    """calc_mmd(kernel, X, Y) -- src/metrics/mmd.jl:23-28; X d x n, Y d x m (columns = samples).
    With torch.distributed initialised the tile grid is dealt across ranks and the three sums are all-reduced."""
    from . import parallel
    ctx = ctx or default_context()
    lam = np.ascontiguousarray(kernel.lengthscales, dtype=np.float64)
    d = len(lam)
    ka, pa, mema, n = _colmajor(X, d)
    kb, pb, memb, m = _colmajor(Y, d)
    if mema != memb:
        raise ValueError("X and Y must live in the same memory space")
    sums = np.zeros(3)
    rank, world = parallel.rank_world()
    if mema == L.DEVICE:
        ctx.order_after_torch()
    ctx._check(ctx.lib.bosip_b200_mmd_sums(ctx.h, kernel.kernel, d, _dptr(lam), pa, n, pb, m, mema, rank, world, _dptr(sums)))
    sums = parallel.allreduce_sum(sums)
    return float(sums[0] / (n * n) + sums[1] / (m * m) - 2.0 * sums[2] / (n * m))


@dataclass
class MMDMetric:
    """src/metrics/mmd.jl:13-15"""
    kernel: Kernel


def _lme(m: float, s: float, G: int) -> float:
    if math.isinf(m) or s <= 0.0:   # empty / all -Inf, or an exactly-zero scaled sum (identical distributions)
        return m if math.isinf(m) else -math.inf
    return m + math.log(s) - math.log(G)


def calc_tv(log_ws, approx_logvals, true_logvals, ctx: Optional[Context] = None) -> float:
    """calc_tv(log_ws, approx_logvals, true_logvals) -- src/metrics/tv.jl:59-73.  Each rank passes its slice of the grid
    when torch.distributed is initialised; the (max, sum) pairs are combined across ranks between the two stages."""
    from . import parallel
    ctx = ctx or default_context()
    kw, pw, mem, G = _colmajor(log_ws); ka, pa, m2, _ = _colmajor(approx_logvals); kt, pt, m3, _ = _colmajor(true_logvals)
    if not (mem == m2 == m3):
        raise ValueError("inputs must live in the same memory space")
    if mem == L.DEVICE:
        ctx.order_after_torch()
    if parallel.rank_world()[1] == 1:      # one GPU: both stages in one call, host vectors uploaded once
        out = np.zeros(1)
        ctx._check(ctx.lib.bosip_b200_tv(ctx.h, pw, pa, pt, G, mem, _dptr(out)))
        return float(out[0])
    s1 = np.zeros(4)
    ctx._check(ctx.lib.bosip_b200_tv_stage1(ctx.h, pw, pa, pt, G, mem, _dptr(s1)))
    (ma, sa), (mt, st_), Gtot = parallel.lse_allgather(s1[0], s1[1]), parallel.lse_allgather(s1[2], s1[3]), parallel.allreduce_count(G)
    eva, evt = _lme(ma, sa, Gtot), _lme(mt, st_, Gtot)
    s2 = np.zeros(2)
    ctx._check(ctx.lib.bosip_b200_tv_stage2(ctx.h, pw, pa, pt, G, mem, eva, evt, _dptr(s2)))
    md, sd = parallel.lse_allgather(s2[0], s2[1])
    return 0.5 * math.exp(_lme(md, sd, Gtot))


class TVMetric:
    """TVMetric(; grid, ws | log_ws, true_logpost | true_logvals) -- src/metrics/tv.jl:21-41"""

    def __init__(self, grid, ws=None, log_ws=None, true_logpost=None, true_logvals=None):
        assert (ws is None) != (log_ws is None)
        self.grid = np.asfortranarray(np.asarray(grid, float))
        self.log_ws = np.log(np.asarray(ws, float)) if log_ws is None else np.asarray(log_ws, float)
        if true_logvals is None and true_logpost is not None:
            true_logvals = np.asarray(true_logpost(self.grid), float)
        self.true_logvals = true_logvals


def calculate_metric(metric, a, b, ctx: Optional[Context] = None) -> float:
    """calculate_metric(::MMDMetric, true_samples, approx_samples) (mmd.jl:17-21) /
       calculate_metric(::TVMetric, true_logpost, approx_logpost) (tv.jl:43-57; closures are called on the whole grid)."""
    if isinstance(metric, MMDMetric):
        return calc_mmd(metric.kernel, a, b, ctx)
    if isinstance(metric, TVMetric):
        true_logvals = metric.true_logvals if metric.true_logvals is not None else np.asarray(a(metric.grid), float)
        approx_logvals = np.asarray(b(metric.grid), float)
        return calc_tv(metric.log_ws, approx_logvals, true_logvals, ctx)
    raise TypeError(type(metric))


# ----------------------------------------------------------------------------------------------- marginal-integration sweep
def generate_LHC(bounds, count: int, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """BOSS.generate_LHC(bounds, count): d x count Latin-hypercube samples (ext/CairoExt.jl:92)."""
    rng = rng or np.random.default_rng()
    lb, ub = np.asarray(bounds[0], float), np.asarray(bounds[1], float)
    d = len(lb)
    u = (np.stack([rng.permutation(count) for _ in range(d)]) + rng.uniform(size=(d, count))) / count
    return lb[:, None] + (ub - lb)[:, None] * u


def normalize_prob_vals(ys: np.ndarray, *steps: float) -> np.ndarray:
    """normalize_prob_vals! (ext/CairoExt.jl:410-430): trapezoid-rule normalisation on the plotting grid."""
    ys = np.array(ys, dtype=float)
    if ys.ndim == 1:
        total = ys.sum() - 0.5 * (ys[0] + ys[-1])
        return np.full_like(ys, 1.0 / (steps[0] * (len(ys) - 1))) if total == 0.0 else ys / (steps[0] * total)
    total = ys.sum() - 0.5 * (ys[0, :].sum() + ys[-1, :].sum() + ys[:, 0].sum() + ys[:, -1].sum())
    total += 0.25 * (ys[0, 0] + ys[0, -1] + ys[-1, 0] + ys[-1, -1])
    if total == 0.0:
        return np.full_like(ys, 1.0 / (steps[0] * steps[1] * (ys.shape[0] - 1) * (ys.shape[1] - 1)))
    return ys / (steps[0] * steps[1] * total)


_FUNC_BITS = {"approx_posterior": L.WANT_APPROX, "posterior_mean": L.WANT_MEAN, "posterior_variance": L.WANT_VAR}


def compute_marginals_int(bosip: BosipProblem, func: str = "posterior_mean", grid_size: int = 200, lhc_grid_size: int = 200,
                          xs=None, plot_diagonal: bool = True, plot_pair_marginals: bool = True, rng=None, normalize: bool = True):
    """compute_marginals_int / _compute_marginals_int (ext/CairoExt.jl:137-246) with the whole sweep on the device: the
    (d*gs + C(d,2)*gs^2) * G query points are generated, evaluated and averaged there; only the means come back.
    Returns {'marginals': {(a, b): {'xs': (...), 'ys': ...}}, 'bounds', 'X', 'lhc'} with 0-based dims (PlotData)."""
    post = bosip.model_posterior()
    if not isinstance(post, B200GPPosterior):
        raise NotImplementedError("the device sweep needs a single conditioned GP (MAP parameters)")
    lb, ub = (np.ascontiguousarray(b, dtype=np.float64) for b in bosip.bounds)
    d = post.d
    if xs is None:
        xs = generate_LHC((lb, ub), lhc_grid_size, rng)
    xs = np.asfortranarray(np.asarray(xs, dtype=np.float64))
    G = xs.shape[1]
    gs = int(grid_size)
    npairs = d * (d - 1) // 2
    diag = np.empty((gs, d), order="F") if plot_diagonal else None
    pairs = np.empty((gs, gs, npairs), order="F") if (plot_pair_marginals and npairs) else None
    lk = bosip.likelihood._c(); pr = bosip.x_prior._c()
    post.ctx._check(post.ctx.lib.bosip_b200_marginals_int(post.h, C.byref(lk), C.byref(pr), _dptr(xs), G, gs, _dptr(lb), _dptr(ub),
                                                         _FUNC_BITS[func], _dptr(diag) if diag is not None else None,
                                                         _dptr(pairs) if pairs is not None else None))
    steps = (ub - lb) / (gs - 1)                                       # plot_steps, ext/CairoExt.jl:432-435
    axes = [np.linspace(lb[k], ub[k], gs) for k in range(d)]
    marg = {}
    if diag is not None:
        for k in range(d):
            ys = diag[:, k]
            marg[(k, k)] = {"xs": (axes[k],), "ys": normalize_prob_vals(ys, steps[k]) if normalize else ys.copy()}
    if pairs is not None:
        pi = 0
        for a in range(d):
            for b in range(a + 1, d):
                ys = pairs[:, :, pi]
                ys = normalize_prob_vals(ys, steps[a], steps[b]) if normalize else ys.copy()
                marg[(a, b)] = {"xs": (axes[a], axes[b]), "ys": ys}
                marg[(b, a)] = {"xs": (axes[b], axes[a]), "ys": ys.T}
                pi += 1
    return {"marginals": marg, "bounds": (lb, ub), "X": bosip.X, "lhc": xs}


# ----------------------------------------------------------------------------------------------- the BO loop (bosip!)
@dataclass
class DataLimit:
    """BOSS.DataLimit(n): stop when the dataset holds n simulations (examples/simple/main.jl:63)."""
    n: int

    def __call__(self, problem: BosipProblem) -> bool:
        return problem.X.shape[1] < self.n


@dataclass
class IterLimit:
    """BOSS.IterLimit(n)."""
    n: int
    _it: int = 0

    def __call__(self, problem: BosipProblem) -> bool:
        self._it += 1
        return self._it <= self.n


def bosip(problem: BosipProblem, simulator: Callable, acquisition=None, acq_maximizer: Optional[B200CandidateAM] = None,
          term_cond=None, model_fitter: Optional[Callable] = None, callback: Optional[Callable] = None, info: bool = False):
    """bosip!(bosip; model_fitter, acq_maximizer, term_cond, options) -- src/main.jl:43-64, the loop BOSS.bo! runs:
    [estimate parameters] -> maximise the acquisition over the candidate set -> evaluate the simulator -> augment the data.

    Only the acquisition step is this repository's business (fused evaluate + argmax on the device); `model_fitter`
    (problem -> GaussianProcess with new hyper-parameters, BOSS's MAP/BI fit) is an optional host callback and defaults to
    keeping the current hyper-parameters."""
    acquisition = acquisition or MaxVar()
    term_cond = term_cond or IterLimit(1)
    if acq_maximizer is None:
        lb, ub = problem.bounds
        acq_maximizer = B200CandidateAM(CandidateSource.philox(0, lb, ub, 100_000))
    it = 0
    while term_cond(problem):
        if model_fitter is not None:
            problem.model = model_fitter(problem)
            problem.augment(np.empty((problem.X.shape[0], 0)), np.empty((problem.Y.shape[0], 0)))  # drop the cached posterior
        x, val = acq_maximizer.maximize_acquisition(problem, acquisition)
        y = np.asarray(simulator(x), dtype=np.float64).reshape(-1)
        problem.augment(x, y)
        it += 1
        if info:
            print(f"[bosip] iter {it}: x = {np.round(x, 4)}, acq = {val:.4g}, data = {problem.X.shape[1]}")
        if callback is not None:
            callback(problem, it, x, val)
    return problem


# ----------------------------------------------------------------------------------------------- multi-GPU behind the C ABI
class _BorrowedContext(Context):
    """A per-device context owned by a MultiContext (never shut down on its own)."""

    def __init__(self, lib, handle, device):   # noqa: super().__init__ would create a new context
        self.lib, self.h, self.device = lib, C.c_void_p(handle), device

    def close(self):
        self.h = None


class MultiContext:
    """bosip_b200_init_multi: ONE host process driving several GPUs through the C ABI -- what a Julia host gets without torchrun
    (include/bosip_b200.h, csrc/multi.cu).  Points are sharded into contiguous ranges, one host thread per device, results
    combined in rank order inside the library.  Large buffers are host arrays."""

    def __init__(self, devices):
        self.lib = L.load()
        devs = list(range(devices)) if isinstance(devices, int) else list(devices)
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        rc = self.lib.bosip_b200_init_multi(len(devs), arr, C.byref(h))
        if rc != 0:
            raise BosipB200Error(rc, (self.lib.bosip_b200_multi_last_error(None) or b"").decode())
        self.h, self.devices = h, devs

    @property
    def size(self) -> int:
        return self.lib.bosip_b200_multi_size(self.h)

    def context(self, rank: int) -> Context:
        return _BorrowedContext(self.lib, self.lib.bosip_b200_multi_context(self.h, rank), self.devices[rank])

    def _check(self, rc: int):
        if rc == 0:
            return
        msg = (self.lib.bosip_b200_multi_last_error(self.h) or b"").decode()
        if rc == L.EDOMAIN:
            raise DomainError(rc, msg)
        if rc == L.ENOTPD:
            raise PosDefException(rc, msg)
        raise BosipB200Error(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.bosip_b200_shutdown_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- metrics (src/metrics/mmd.jl:23-28, tv.jl:59-73)
    def calc_mmd(self, kernel: "Kernel", X, Y) -> float:
        lam = np.ascontiguousarray(kernel.lengthscales, dtype=np.float64)
        ka, pa, ma, n = _colmajor(np.asarray(X), len(lam)); kb, pb, mb, m = _colmajor(np.asarray(Y), len(lam))
        out = C.c_double()
        self._check(self.lib.bosip_b200_mmd_multi(self.h, kernel.kernel, len(lam), _dptr(lam), pa, n, pb, m, C.byref(out)))
        return out.value

    def calc_tv(self, log_ws, approx_logvals, true_logvals) -> float:
        kw, pw, _, G = _colmajor(np.asarray(log_ws)); ka, pa, _, _ = _colmajor(np.asarray(approx_logvals)); kt, pt, _, _ = _colmajor(np.asarray(true_logvals))
        out = C.c_double()
        self._check(self.lib.bosip_b200_tv_multi(self.h, pw, pa, pt, G, C.byref(out)))
        return out.value


class MultiGPPosterior:
    """N replicas of one conditioned GP (bosip_b200_gp_fit_multi)."""

    def __init__(self, mctx: MultiContext, handle, model: GaussianProcess, d: int, N: int, p: int):
        self.mctx, self.h, self.model, self.d, self.N, self.p = mctx, handle, model, d, N, p

    def free(self):
        if self.h:
            self.mctx.lib.bosip_b200_gp_free_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def model_posterior_multi(model: GaussianProcess, X, Y, mctx: MultiContext) -> MultiGPPosterior:
    X = np.asfortranarray(np.asarray(X, dtype=np.float64)); Y = np.asfortranarray(np.asarray(Y, dtype=np.float64))
    d, N = X.shape; p = Y.shape[0]
    lam = np.asfortranarray(np.asarray(model.lengthscales, dtype=np.float64).reshape(d, p))
    alpha = np.ascontiguousarray(model.amplitudes, dtype=np.float64); sigma = np.ascontiguousarray(model.noise_std, dtype=np.float64)
    if model.mean is not None:
        raise NotImplementedError("the multi-GPU entry points take a zero prior mean")
    h = C.c_void_p(); o = _opts(model)
    mctx._check(mctx.lib.bosip_b200_gp_fit_multi(mctx.h, model.kernel, d, N, p, _dptr(X), _dptr(Y), _dptr(lam), _dptr(alpha), _dptr(sigma),
                                                 None, C.byref(o), C.byref(h)))
    return MultiGPPosterior(mctx, h, model, d, N, p)


def posterior_eval_multi(post: MultiGPPosterior, like, prior: Optional[ProductPrior], X, want: int, logprior=None, acq=None,
                         index_offset: int = 0, out=None):
    """posterior_eval through bosip_b200_posterior_eval_multi: host X (d x M) sharded over the devices of `post.mctx`."""
    keep, addr, mem, M = _colmajor(np.asarray(X), post.d)
    lk = like._c()
    lp_keep = lp_addr = None
    if logprior is not None:
        lp_keep, lp_addr, _, _ = _colmajor(np.asarray(logprior))
    pr = prior._c(lp_addr) if prior is not None else L.Prior(post.d, None, None, lp_addr)
    outs, ptrs = {}, []
    for bit, name in ((L.WANT_APPROX, "log_approx_post"), (L.WANT_MEAN, "log_post_mean"), (L.WANT_VAR, "log_post_var")):
        if want & bit:
            o = out[name] if (out is not None and name in out) else np.empty(M)
            outs[name] = o; ptrs.append(o.ctypes.data)
        else:
            ptrs.append(None)
    ncl = C.c_int64(0); bi = C.c_int64(-1); bv = C.c_double(0.0)
    post.mctx._check(post.mctx.lib.bosip_b200_posterior_eval_multi(
        post.h, C.byref(lk), C.byref(pr), addr, M, None, want, ptrs[0], ptrs[1], ptrs[2], C.byref(ncl),
        acq.acq_id if acq is not None else -1, index_offset, C.byref(bi), C.byref(bv)))
    outs["n_clamped"] = ncl.value
    if acq is not None:
        outs["best_idx"], outs["best_val"] = bi.value, bv.value
    return outs


def acq_argmax_multi(post: MultiGPPosterior, like, prior, acq, src: CandidateSource, start: int = 0, count: Optional[int] = None):
    """Fused evaluate + argmax over candidates start..start+count-1 of `src`, sharded over the devices inside the library."""
    count = src.count - start if count is None else count
    cs = L.CandSrc(); cs.kind = src.kind; keep = None
    if src.kind == L.CAND_POINTS:
        keep, addr, mem, _ = _colmajor(np.asarray(src.points)[:, start:start + count], post.d)
        cs.points = addr; cs.mem = L.HOST
    else:
        cs.seed = src.seed; cs.lb = _dptr(src.lb); cs.ub = _dptr(src.ub); cs.n_per_dim = src.n_per_dim
    lk = like._c(); pr = prior._c() if prior is not None else L.Prior(post.d, None, None, None)
    bi, bv = C.c_int64(-1), C.c_double()
    bx = np.full(post.d, np.nan)
    post.mctx._check(post.mctx.lib.bosip_b200_acq_argmax_multi(post.h, C.byref(lk), C.byref(pr), C.byref(cs), count, start, acq.acq_id,
                                                               C.byref(bi), C.byref(bv), _dptr(bx)))
    return bi.value, bv.value, bx


def evidence_multi(post: MultiGPPosterior, like, prior: ProductPrior, xs, which: str = "approx") -> float:
    keep, addr, mem, S = _colmajor(np.asarray(xs), post.d)
    lk = like._c(); pr = prior._c()
    out = C.c_double()
    post.mctx._check(post.mctx.lib.bosip_b200_evidence_sum_multi(post.h, C.byref(lk), C.byref(pr), addr, S,
                                                                 L.WANT_APPROX if which == "approx" else L.WANT_MEAN, C.byref(out)))
    return out.value / S
