"""bosip_b200 -- B200-native evaluator for BOSIP.jl's batched GP-posterior / likelihood / acquisition / metric path.

The directory is named `bosip.jl_b200` (not importable by that name); import it as `bosip_b200` through the shim at the
repository root.  All compute goes through libbosip_b200.so (hand-written sm_100a CUDA, C ABI in include/bosip_b200.h);
importing the package does not need a GPU, creating a Context does."""
from . import _lib, parallel, synth  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (B200CandidateAM, B200GPPosterior, BosipB200Error, BosipProblem, CandidateSource, Context, DomainError,
                  GaussianLikelihood, GaussianProcess, Kernel, LogMaxVar, MaxVar, MMDMetric, NormalLikelihood,
                  PosDefException, ProductPrior, TVMetric, acq_argmax, approx_likelihood, approx_posterior, calc_mmd,
                  calc_tv, calculate_metric, candidates, default_context, evidence, likelihood_mean, likelihood_variance,
                  log_approx_likelihood, log_approx_posterior, log_likelihood_mean, log_likelihood_variance,
                  log_posterior_mean, log_posterior_variance, model_posterior, model_posterior_from_factors,
                  posterior_eval, posterior_mean, posterior_variance, normal_likelihood_eval, loglike, loglike_marginal,
                  log_approx_marginal_likelihood, log_marginal_likelihood_mean, log_sq_likelihood_mean, compute_marginals_int,
                  generate_LHC, normalize_prob_vals, MultiContext, MultiGPPosterior, model_posterior_multi, posterior_eval_multi,
                  acq_argmax_multi, evidence_multi, LogNormalLikelihood, NormalDiffLikelihood, NormalSumLikelihood, ExpLikelihood, bosip, DataLimit, IterLimit)

from .sampling import (AEConfidence, AMIS, AMISSampler, AnalyticalFitter, NormalProposal, OptMMDMetric, amis_log_weights,  # noqa: F401,E402
                       approx_cutoff_area, confidence_iou, exp_weights, find_cutoff, mixture_pdf_sum, resample, sample_posterior,
                       sample_posterior_pure, set_iou, weighted_moments, weighted_quantile)

__version__ = "0.1.0"
