# BosipB200.jl -- Julia glue that makes libbosip_b200.so a drop-in for BOSIP.jl's batched evaluation path.
#
# NOT RUN IN THIS REPOSITORY'S CI: the build image has no Julia toolchain (SURVEY.md section 0), so this file is the
# reference-side binding a maintainer would add; the identical C ABI is exercised by the Python/ctypes harness
# (bosip.jl_b200/api.py, tests/).  Seams (SURVEY.md section 8b):
#   1. B200GPPosterior <: BOSS.ModelPosterior        -> mean/var/mean_and_var (vector -> p, matrix -> p x M)
#   2. fused Likelihood methods for NormalLikelihood  -> one ccall per X instead of three GP predicts
#   3. B200CandidateAM <: BOSS.AcquisitionMaximizer   -> fused evaluate + argmax over a candidate set
#   4. calc_mmd / calc_tv methods                      -> metrics
module BosipB200

using BOSIP, BOSS, Distributions, LinearAlgebra

const LIB = get(ENV, "BOSIP_B200_LIB", joinpath(@__DIR__, "..", "bosip.jl_b200", "libbosip_b200.so"))

const HOST, DEVICE = Cint(0), Cint(1)
const WANT_APPROX, WANT_MEAN, WANT_VAR = Cuint(1), Cuint(2), Cuint(4)
const EDOMAIN, ENOTPD = Cint(-4), Cint(-3)

struct GpOpts;  jitter::Cdouble; pred_noise::Int32; clamp_var::Int32; end
struct LikeNormal; p::Int32; kind::Int32; z_obs::Ptr{Cdouble}; std_obs::Ptr{Cdouble}; sum_lengths::Ptr{Int32}; end  # kind: 0 Normal, 1 LogNormal, 2 NormalDiff, 3 NormalSum, 4 Exp
struct Prior; d::Int32; kind::Ptr{Int32}; params::Ptr{Cdouble}; logprior::Ptr{Cdouble}; end
struct CandSrc; kind::Int32; mem::Int32; points::Ptr{Cdouble}; seed::UInt64; lb::Ptr{Cdouble}; ub::Ptr{Cdouble}; n_per_dim::Int64; end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bosip_b200_init, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error(unsafe_string(ccall((:bosip_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> ccall((:bosip_b200_shutdown, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    end
end
const CTX = Ref{Context}()
context() = (isassigned(CTX) || (CTX[] = Context(parse(Int, get(ENV, "LOCAL_RANK", "0")))); CTX[])

function check(ctx::Context, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:bosip_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == EDOMAIN && throw(DomainError(NaN, msg))            # src/posterior/likelihood.jl:61-65
    rc == ENOTPD && throw(LinearAlgebra.PosDefException(1))  # Julia's cholesky
    error("bosip_b200 [$rc]: $msg")
end

kernel_id(::BOSS.GaussianKernel) = Cint(0); kernel_id(::BOSS.Matern32Kernel) = Cint(1); kernel_id(::BOSS.Matern52Kernel) = Cint(2)

# ---- seam 1: ModelPosterior -------------------------------------------------------------------------------------
mutable struct B200GPPosterior{M} <: BOSS.ModelPosterior{M}
    ctx::Context; h::Ptr{Cvoid}; d::Int; p::Int
end

"Wraps a BOSS.GaussianProcess so that `BOSS.model_posterior(problem)` (src/posterior/log_posterior.jl:81) returns the device-backed posterior."
struct B200GaussianProcess{G<:BOSS.GaussianProcess} <: BOSS.SurrogateModel; gp::G; end

function BOSS.model_posterior(model::B200GaussianProcess, params::BOSS.UniFittedParams, data::BOSS.ExperimentData)
    ctx = context()
    X, Y = Matrix{Float64}(data.X), Matrix{Float64}(data.Y)          # d x N, p x N (src/subset_problem.jl:46-50)
    λ, α, σ = params.lengthscales, params.amplitude, params.noise_std  # d x p, p, p (src/subset_problem.jl:38-44)
    d, N = size(X); p = size(Y, 1)
    mX = isnothing(model.gp.mean) ? C_NULL : Matrix{Float64}(hcat(model.gp.mean.(eachcol(X))...))
    h = Ref{Ptr{Cvoid}}(C_NULL); opts = Ref(GpOpts(0.0, 0, 1))
    GC.@preserve X Y λ α σ mX begin
        check(ctx, ccall((:bosip_b200_gp_fit, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{GpOpts}, Ref{Ptr{Cvoid}}),
            ctx.h, kernel_id(model.gp.kernel), d, N, p, X, Y, λ, α, σ, mX, opts, h))
    end
    post = B200GPPosterior{typeof(model)}(ctx, h[], d, p)
    finalizer(q -> ccall((:bosip_b200_gp_free, LIB), Cint, (Ptr{Cvoid},), q.h), post)
end

function _predict(post::B200GPPosterior, X::AbstractMatrix{<:Real}; want_mu=true, want_var=true)
    Xc = Matrix{Float64}(X); M = size(Xc, 2)
    mu = want_mu ? Matrix{Float64}(undef, post.p, M) : C_NULL      # p x M (test/unit/utils.jl:8-11)
    var = want_var ? Matrix{Float64}(undef, post.p, M) : C_NULL
    GC.@preserve Xc mu var check(post.ctx, ccall((:bosip_b200_gp_predict, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), post.h, Xc, M, HOST, C_NULL, mu, var))
    return mu, var
end
BOSS.mean(post::B200GPPosterior, X::AbstractMatrix{<:Real}) = _predict(post, X; want_var=false)[1]
BOSS.var(post::B200GPPosterior, X::AbstractMatrix{<:Real}) = _predict(post, X; want_mu=false)[2]
BOSS.mean_and_var(post::B200GPPosterior, X::AbstractMatrix{<:Real}) = _predict(post, X)
BOSS.mean(post::B200GPPosterior, x::AbstractVector{<:Real}) = vec(BOSS.mean(post, reshape(x, :, 1)))
BOSS.var(post::B200GPPosterior, x::AbstractVector{<:Real}) = vec(BOSS.var(post, reshape(x, :, 1)))
BOSS.mean_and_var(post::B200GPPosterior, x::AbstractVector{<:Real}) = vec.(_predict(post, reshape(x, :, 1)))

# ---- seam 2: fused closures -------------------------------------------------------------------------------------
function _prior_desc(prior::Product)   # Product of Uniform / Normal / truncated(Normal); anything else -> host logpdf values
    kinds = Int32[]; prm = Float64[]
    for c in prior.v
        if c isa Uniform;          push!(kinds, 0); append!(prm, (c.a, c.b, 0.0, 0.0))
        elseif c isa Normal;       push!(kinds, 1); append!(prm, (c.μ, c.σ, 0.0, 0.0))
        elseif c isa Truncated{<:Normal}; push!(kinds, 2); append!(prm, (c.untruncated.μ, c.untruncated.σ, c.lower, c.upper))
        else return nothing end
    end
    return kinds, prm
end

# Likelihood descriptors of the closed-form Gaussian family (include/bosip_b200.h, BOSIP_B200_LIKE_*): (kind, z, so, sum_lengths).
# The returned arrays must stay alive (GC.@preserve) while the struct built from their pointers is in use.
_like_desc(l::NormalLikelihood)     = (Int32(0), Vector{Float64}(l.z_obs), Vector{Float64}(l.std_obs), Int32[])       # src/likelihoods/normal.jl:16-19
_like_desc(l::LogNormalLikelihood)  = (Int32(1), Vector{Float64}(l.log_z_obs), Vector{Float64}(l.CV), Int32[])         # src/likelihoods/lognormal.jl:26-42
_like_desc(l::NormalDiffLikelihood) = (Int32(2), Float64[], Vector{Float64}(l.std_obs), Int32[])                       # src/likelihoods/normal_diff.jl:17-19
_like_desc(l::NormalSumLikelihood)  = (Int32(3), Vector{Float64}(l.z_obs), Vector{Float64}(l.std_obs), Vector{Int32}(l.sum_lengths))  # normal_sum.jl:18-27
_like_desc(l::ExpLikelihood)        = (Int32(4), Float64[], Float64[], Int32[])                                         # src/likelihoods/exp.jl:7
const B200Likelihood = Union{NormalLikelihood, LogNormalLikelihood, NormalDiffLikelihood, NormalSumLikelihood, ExpLikelihood}
_obs_dim(kind, z, so, sl) = kind == 4 ? 1 : length(so)
_like_struct(kind, z, so, sl) = LikeNormal(_obs_dim(kind, z, so, sl), kind, isempty(z) ? C_NULL : pointer(z), isempty(so) ? C_NULL : pointer(so),
                                           isempty(sl) ? C_NULL : pointer(sl))

function _posterior_eval(post::B200GPPosterior, like::NormalLikelihood, x_prior, X::AbstractMatrix{<:Real}, want::Cuint)
    Xc = Matrix{Float64}(X); M = size(Xc, 2)
    out = [((want & b) != 0) ? Vector{Float64}(undef, M) : C_NULL for b in (WANT_APPROX, WANT_MEAN, WANT_VAR)]
    desc = isnothing(x_prior) ? (Int32[], Float64[]) : _prior_desc(x_prior)
    lp = (isnothing(desc) && !isnothing(x_prior)) ? logpdf.(Ref(x_prior), eachcol(Xc)) : C_NULL   # src/posterior/log_posterior.jl:228-229
    kinds, prm = isnothing(desc) ? (Int32[], Float64[]) : desc
    z, so = like.z_obs, like.std_obs
    ncl = Ref{Int64}(0)
    GC.@preserve Xc out kinds prm lp z so begin
        lk = Ref(LikeNormal(length(z), 0, pointer(z), pointer(so), C_NULL))
        pr = Ref(Prior(post.d, isempty(kinds) ? C_NULL : pointer(kinds), isempty(prm) ? C_NULL : pointer(prm), lp === C_NULL ? C_NULL : pointer(lp)))
        check(post.ctx, ccall((:bosip_b200_posterior_eval, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Cuint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Int64}),
            post.h, lk, pr, Xc, M, HOST, C_NULL, want, out[1], out[2], out[3], ncl))
    end
    return out
end

_closure(post, like, prior, want, i) = (f(X::AbstractMatrix{<:Real}) = _posterior_eval(post, like, prior, X, want)[i];
                                        f(x::AbstractVector{<:Real}) = f(reshape(x, :, 1))[1]; f)

# more specific methods than the defaults of src/posterior/likelihood.jl:19-59
BOSIP.log_approx_likelihood(like::NormalLikelihood, post::B200GPPosterior) = _closure(post, like, nothing, WANT_APPROX, 1)
BOSIP.log_likelihood_mean(like::NormalLikelihood, post::B200GPPosterior) = _closure(post, like, nothing, WANT_MEAN, 2)
BOSIP.log_likelihood_variance(like::NormalLikelihood, post::B200GPPosterior) = _closure(post, like, nothing, WANT_VAR, 3)

# posterior-level closures (src/posterior/log_posterior.jl:15-61) with the prior fused in as well
_is_b200(bosip) = bosip.problem.model isa B200GaussianProcess && bosip.likelihood isa NormalLikelihood
function fused_log_posterior_variance(bosip::BosipProblem)
    @assert _is_b200(bosip)
    _closure(BOSS.model_posterior(bosip.problem), bosip.likelihood, bosip.x_prior, WANT_VAR, 3)
end
fused_log_posterior_mean(bosip) = _closure(BOSS.model_posterior(bosip.problem), bosip.likelihood, bosip.x_prior, WANT_MEAN, 2)
fused_log_approx_posterior(bosip) = _closure(BOSS.model_posterior(bosip.problem), bosip.likelihood, bosip.x_prior, WANT_APPROX, 1)

# ---- seam 3: acquisition maximiser ------------------------------------------------------------------------------
"Candidate-set maximiser (SamplingAM/GridAM semantics, docs/src/hyperparams.md:69): `candidates` d x M, or Philox U(domain)."
struct B200CandidateAM <: BOSS.AcquisitionMaximizer
    candidates::Union{Matrix{Float64}, Nothing}; seed::UInt64; count::Int64
end
function BOSS.maximize_acquisition(am::B200CandidateAM, problem::BOSS.BossProblem, options::BOSS.BossOptions)
    wrapper = problem.acquisition::BOSIP.AcqWrapper                  # src/types/acquisition.jl:24-30
    acq = wrapper.acq isa MaxVar ? Cint(0) : wrapper.acq isa LogMaxVar ? Cint(1) : error("B200CandidateAM supports MaxVar / LogMaxVar")
    bosip = wrapper.bosip; post = BOSS.model_posterior(problem); like = bosip.likelihood
    kinds, prm = _prior_desc(bosip.x_prior); z, so = like.z_obs, like.std_obs
    lb, ub = problem.domain.bounds
    best_idx, best_val, best_x = Ref{Int64}(-1), Ref{Cdouble}(0.0), Vector{Float64}(undef, post.d)
    GC.@preserve am kinds prm z so lb ub best_x begin
        src = isnothing(am.candidates) ? CandSrc(1, 0, C_NULL, am.seed, pointer(lb), pointer(ub), 0) :
                                         CandSrc(0, 0, pointer(am.candidates), 0, C_NULL, C_NULL, 0)
        M = isnothing(am.candidates) ? am.count : size(am.candidates, 2)
        check(post.ctx, ccall((:bosip_b200_acq_argmax, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ref{CandSrc}, Int64, Int64, Cint, Ref{Int64}, Ref{Cdouble}, Ptr{Cdouble}),
            post.h, Ref(LikeNormal(length(z), 0, pointer(z), pointer(so), C_NULL)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)),
            Ref(src), M, 0, acq, best_idx, best_val, best_x))
    end
    return best_x, best_val[]   # BOSS 0.6 return convention (x vs (x, val)) to be confirmed -- SURVEY.md Appendix B item 6
end

# ---- seam 4: metrics --------------------------------------------------------------------------------------------
function BOSIP.calc_tv(log_ws::Vector{Float64}, approx_logvals::Vector{Float64}, true_logvals::Vector{Float64})   # src/metrics/tv.jl:59-73
    ctx = context(); out = Ref{Cdouble}(0.0)
    check(ctx, ccall((:bosip_b200_tv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Ref{Cdouble}),
        ctx.h, log_ws, approx_logvals, true_logvals, length(log_ws), HOST, out))
    return out[]
end
"calc_mmd for ARD SE / Matern kernels (src/metrics/mmd.jl:23-28); X, Y as column matrices."
function calc_mmd_b200(kernel_id::Integer, lengthscales::Vector{Float64}, X::Matrix{Float64}, Y::Matrix{Float64})
    ctx = context(); out = Ref{Cdouble}(0.0)
    check(ctx, ccall((:bosip_b200_mmd, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Cint, Ref{Cdouble}),
        ctx.h, kernel_id, size(X, 1), lengthscales, X, size(X, 2), Y, size(Y, 2), HOST, out))
    return out[]
end


# ---- the other closed-form likelihoods (src/likelihoods/lognormal.jl:53-91, exp.jl:19-63, normal_sum.jl:68-115, normal_diff.jl:28-73),
#      Bayesian-inference ensembles (src/posterior/log_posterior.jl:86-95,120-129,153-162) and evidence (src/posterior/evidence.jl:19-24)
"One fused sweep for any closed-form likelihood and one or several conditioned GPs (`posts`: MultiFittedParams -> a vector)."
function posterior_eval_b200(posts::Vector{<:B200GPPosterior}, like::B200Likelihood, x_prior, X::AbstractMatrix{<:Real}, want::Cuint;
                             acq::Integer=-1, index_offset::Integer=0)
    post = posts[1]; Xc = Matrix{Float64}(X); M = size(Xc, 2)
    out = [((want & b) != 0) ? Vector{Float64}(undef, M) : C_NULL for b in (WANT_APPROX, WANT_MEAN, WANT_VAR)]
    desc = isnothing(x_prior) ? (Int32[], Float64[]) : _prior_desc(x_prior)
    lp = (isnothing(desc) && !isnothing(x_prior)) ? logpdf.(Ref(x_prior), eachcol(Xc)) : C_NULL
    kinds, prm = isnothing(desc) ? (Int32[], Float64[]) : desc
    kind, z, so, sl = _like_desc(like)
    hs = Ptr{Cvoid}[q.h for q in posts]
    ncl = Ref{Int64}(0); best_idx = Ref{Int64}(-1); best_val = Ref{Cdouble}(0.0)
    GC.@preserve Xc out kinds prm lp z so sl hs posts begin
        lk = Ref(_like_struct(kind, z, so, sl))
        pr = Ref(Prior(post.d, isempty(kinds) ? C_NULL : pointer(kinds), isempty(prm) ? C_NULL : pointer(prm), lp === C_NULL ? C_NULL : pointer(lp)))
        check(post.ctx, ccall((:bosip_b200_posterior_eval_bi, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Cuint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
             Ref{Int64}, Cint, Int64, Ref{Int64}, Ref{Cdouble}),
            hs, length(hs), lk, pr, Xc, M, HOST, C_NULL, want, out[1], out[2], out[3], ncl, acq, index_offset, best_idx, best_val))
    end
    return (log_approx_post=out[1], log_post_mean=out[2], log_post_var=out[3], n_clamped=ncl[], best_idx=best_idx[], best_val=best_val[])
end
posterior_eval_b200(post::B200GPPosterior, args...; kw...) = posterior_eval_b200([post], args...; kw...)

# more specific methods for the remaining closed-form likelihoods and for MultiFittedParams (a vector of posteriors)
for L in (:LogNormalLikelihood, :NormalDiffLikelihood, :NormalSumLikelihood, :ExpLikelihood)
    @eval begin
        BOSIP.log_approx_likelihood(like::$L, post::B200GPPosterior) = _closure_any([post], like, nothing, WANT_APPROX, :log_approx_post)
        BOSIP.log_likelihood_mean(like::$L, post::B200GPPosterior) = _closure_any([post], like, nothing, WANT_MEAN, :log_post_mean)
        BOSIP.log_likelihood_variance(like::$L, post::B200GPPosterior) = _closure_any([post], like, nothing, WANT_VAR, :log_post_var)
    end
end
_closure_any(posts, like, prior, want, key) = (f(X::AbstractMatrix{<:Real}) = getfield(posterior_eval_b200(posts, like, prior, X, want), key);
                                               f(x::AbstractVector{<:Real}) = f(reshape(x, :, 1))[1]; f)
"BI variants of the posterior closures: `posts = BOSS.model_posterior(problem)` with MultiFittedParams (log_posterior.jl:86-95,120-129,153-162)."
fused_log_approx_posterior(bosip::BosipProblem, posts::Vector{<:B200GPPosterior}) = _closure_any(posts, bosip.likelihood, bosip.x_prior, WANT_APPROX, :log_approx_post)
fused_log_posterior_mean(bosip::BosipProblem, posts::Vector{<:B200GPPosterior}) = _closure_any(posts, bosip.likelihood, bosip.x_prior, WANT_MEAN, :log_post_mean)
fused_log_posterior_variance(bosip::BosipProblem, posts::Vector{<:B200GPPosterior}) = _closure_any(posts, bosip.likelihood, bosip.x_prior, WANT_VAR, :log_post_var)

"evidence(post, x_prior; xs) (src/posterior/evidence.jl:19-24) in one sweep; `which` = WANT_APPROX or WANT_MEAN; ensembles are averaged in linear space."
function evidence_b200(posts::Vector{<:B200GPPosterior}, like::B200Likelihood, x_prior::Product, xs::AbstractMatrix{<:Real}; which::Cuint=WANT_MEAN)
    post = posts[1]; Xc = Matrix{Float64}(xs); S = size(Xc, 2)
    kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    hs = Ptr{Cvoid}[q.h for q in posts]; out = Ref{Cdouble}(0.0)
    GC.@preserve Xc kinds prm z so sl hs posts begin
        check(post.ctx, ccall((:bosip_b200_evidence_sum_bi, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Cint, Cuint, Ref{Cdouble}),
            hs, length(hs), Ref(_like_struct(kind, z, so, sl)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)), Xc, S, HOST, which, out))
    end
    return out[] / S
end

"Candidate-set argmax for an ensemble (MultiFittedParams): bosip_b200_acq_argmax_bi."
function acq_argmax_b200(posts::Vector{<:B200GPPosterior}, like::B200Likelihood, x_prior::Product, src::CandSrc, M::Integer, acq::Integer; keep=nothing)
    post = posts[1]; kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    hs = Ptr{Cvoid}[q.h for q in posts]
    best_idx, best_val, best_x = Ref{Int64}(-1), Ref{Cdouble}(0.0), Vector{Float64}(undef, post.d)
    GC.@preserve kinds prm z so sl hs posts keep best_x begin
        check(post.ctx, ccall((:bosip_b200_acq_argmax_bi, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Ref{LikeNormal}, Ref{Prior}, Ref{CandSrc}, Int64, Int64, Cint, Ref{Int64}, Ref{Cdouble}, Ptr{Cdouble}),
            hs, length(hs), Ref(_like_struct(kind, z, so, sl)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)), Ref(src), M, 0, acq,
            best_idx, best_val, best_x))
    end
    return best_x, best_val[], best_idx[]
end

# ---- marginal-integration sweep: the body of `_compute_marginals_int` (ext/CairoExt.jl:146-246) in ONE ccall -------------------
"""
    diag, pairs = marginals_int_b200(post, like, x_prior, lhc, grid_size, lb, ub; which=WANT_MEAN)

`lhc` is the d x G Latin-hypercube matrix `generate_LHC(bounds, lhc_grid_size)` (ext/CairoExt.jl:92).  Returns the UN-normalised means
`diag[i, k]` (grid_size x d) and `pairs[ia, ib, pair]` (pairs ordered (1,2),(1,3),...,(d-1,d) as the reference's loops); apply
`normalize_prob_vals!` (ext/CairoExt.jl:410-430) afterwards.
"""
function marginals_int_b200(post::B200GPPosterior, like::B200Likelihood, x_prior::Product, lhc::AbstractMatrix{<:Real}, grid_size::Integer,
                            lb::Vector{Float64}, ub::Vector{Float64}; which::Cuint=WANT_MEAN)
    d = post.d; G = size(lhc, 2); L = Matrix{Float64}(lhc)
    diag = Matrix{Float64}(undef, grid_size, d); pairs = Array{Float64}(undef, grid_size, grid_size, d * (d - 1) ÷ 2)
    kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    GC.@preserve L diag pairs kinds prm z so sl lb ub begin
        check(post.ctx, ccall((:bosip_b200_marginals_int, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cuint, Ptr{Cdouble}, Ptr{Cdouble}),
            post.h, Ref(_like_struct(kind, z, so, sl)), Ref(Prior(d, pointer(kinds), pointer(prm), C_NULL)), L, G, grid_size, lb, ub, which,
            diag, isempty(pairs) ? C_NULL : pairs))
    end
    return diag, pairs
end

# ---- confidence sets / AEConfidence (src/utils/confidence_set.jl:19-81, src/term_conds/ae_confidence.jl:59-79) ---------------------
"quantile(vals, Distributions.weights(ws), p) on the device (what find_cutoff computes, confidence_set.jl:24-26)."
function weighted_quantile_b200(vals::Vector{Float64}, ws::Vector{Float64}, p::Real)
    ctx = context(); out = Ref{Cdouble}(0.0)
    check(ctx, ccall((:bosip_b200_weighted_quantile, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cdouble, Ref{Cdouble}),
        ctx.h, vals, ws, length(vals), HOST, p, out))
    return out[]
end
BOSIP.find_cutoff(vals::Vector{Float64}, ws::Vector{Float64}, q::Real, ::Val{:b200}) = weighted_quantile_b200(vals, ws, 1.0 - q)
function approx_cutoff_area_b200(vals::Vector{Float64}, ws::Vector{Float64}, c::Real)          # confidence_set.jl:52-56
    ctx = context(); out = Ref{Cdouble}(0.0)
    check(ctx, ccall((:bosip_b200_cutoff_area, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cdouble, Ref{Cdouble}),
        ctx.h, vals, ws, length(vals), HOST, c, out))
    return out[]
end
function set_iou_b200(in_A::AbstractVector{Bool}, in_B::AbstractVector{Bool}, xs_logpdf::Vector{Float64})   # confidence_set.jl:73-81
    ctx = context(); out = Ref{Cdouble}(0.0); a = Vector{Bool}(in_A); b = Vector{Bool}(in_B)              # BitVector -> one byte per flag
    check(ctx, ccall((:bosip_b200_set_iou, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Cdouble}, Int64, Cint, Ref{Cdouble}),
        ctx.h, a, b, xs_logpdf, length(a), HOST, out))
    return out[]
end
"calculate(::AEConfidence, bosip) (ae_confidence.jl:59-79): ONE fused sweep for both closures + one device reduction."
function BOSIP.calculate(cond::BOSIP.AEConfidence, bosip::BosipProblem, ::Val{:b200})
    xs = isnothing(cond.xs) ? rand(bosip.x_prior, cond.samples) : cond.xs
    post = BOSS.model_posterior(bosip.problem)
    r = posterior_eval_b200(post isa Vector ? post : [post], bosip.likelihood, bosip.x_prior, xs, WANT_APPROX | WANT_MEAN)
    xs_logpdf = logpdf.(Ref(bosip.x_prior), eachcol(xs))
    ctx = context(); iou = Ref{Cdouble}(0.0); ca = Ref{Cdouble}(0.0); ce = Ref{Cdouble}(0.0)
    check(ctx, ccall((:bosip_b200_confidence_iou, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
        ctx.h, r.log_approx_post, r.log_post_mean, xs_logpdf, length(xs_logpdf), HOST, cond.q, iou, ca, ce))
    return iou[]
end

# ---- multi-GPU from ONE Julia process (csrc/multi.cu): bosip_b200_init_multi + the `_multi` entry points ----------------------------
mutable struct MultiContext
    h::Ptr{Cvoid}; ndev::Int
    function MultiContext(devices::Vector{<:Integer})
        h = Ref{Ptr{Cvoid}}(C_NULL); devs = Vector{Cint}(devices)
        rc = ccall((:bosip_b200_init_multi, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(devs), devs, h)
        rc == 0 || error(unsafe_string(ccall((:bosip_b200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(h[], length(devs))
        finalizer(c -> ccall((:bosip_b200_shutdown_multi, LIB), Cint, (Ptr{Cvoid},), c.h), m)
    end
end
MultiContext(ndev::Integer) = MultiContext(collect(0:ndev-1))
function check(m::MultiContext, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:bosip_b200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m.h))
    rc == EDOMAIN && throw(DomainError(NaN, msg)); rc == ENOTPD && throw(LinearAlgebra.PosDefException(1))
    error("bosip_b200 [$rc]: $msg")
end
mutable struct B200MultiGPPosterior{M} <: BOSS.ModelPosterior{M}
    m::MultiContext; h::Ptr{Cvoid}; d::Int; p::Int
end
"Replicated fit on every device of `m` (the `_multi` counterpart of BOSS.model_posterior, src/posterior/log_posterior.jl:81)."
function model_posterior_multi(m::MultiContext, model::B200GaussianProcess, params::BOSS.UniFittedParams, data::BOSS.ExperimentData)
    X, Y = Matrix{Float64}(data.X), Matrix{Float64}(data.Y)
    λ, α, σ = params.lengthscales, params.amplitude, params.noise_std
    d, N = size(X); p = size(Y, 1)
    h = Ref{Ptr{Cvoid}}(C_NULL); opts = Ref(GpOpts(0.0, 0, 1))
    GC.@preserve X Y λ α σ begin
        check(m, ccall((:bosip_b200_gp_fit_multi, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{GpOpts}, Ref{Ptr{Cvoid}}),
            m.h, kernel_id(model.gp.kernel), d, N, p, X, Y, λ, α, σ, C_NULL, opts, h))
    end
    post = B200MultiGPPosterior{typeof(model)}(m, h[], d, p)
    finalizer(q -> ccall((:bosip_b200_gp_free_multi, LIB), Cint, (Ptr{Cvoid},), q.h), post)
end
"Posterior closures' sweep over contiguous ranges of X on all devices; same outputs as `posterior_eval_b200`."
function posterior_eval_multi(post::B200MultiGPPosterior, like::B200Likelihood, x_prior::Product, X::AbstractMatrix{<:Real}, want::Cuint;
                              acq::Integer=-1, index_offset::Integer=0)
    Xc = Matrix{Float64}(X); M = size(Xc, 2)
    out = [((want & b) != 0) ? Vector{Float64}(undef, M) : C_NULL for b in (WANT_APPROX, WANT_MEAN, WANT_VAR)]
    kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    ncl = Ref{Int64}(0); best_idx = Ref{Int64}(-1); best_val = Ref{Cdouble}(0.0)
    GC.@preserve Xc out kinds prm z so sl begin
        check(post.m, ccall((:bosip_b200_posterior_eval_multi, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Cuint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Int64},
             Cint, Int64, Ref{Int64}, Ref{Cdouble}),
            post.h, Ref(_like_struct(kind, z, so, sl)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)), Xc, M, C_NULL, want,
            out[1], out[2], out[3], ncl, acq, index_offset, best_idx, best_val))
    end
    return (log_approx_post=out[1], log_post_mean=out[2], log_post_var=out[3], n_clamped=ncl[], best_idx=best_idx[], best_val=best_val[])
end
"B200CandidateAM over all devices: the candidate range is sharded inside the library, the winners combined in rank order."
function maximize_acquisition_multi(am::B200CandidateAM, post::B200MultiGPPosterior, like::B200Likelihood, x_prior::Product, lb::Vector{Float64},
                                    ub::Vector{Float64}, acq::Integer)
    kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    best_idx, best_val, best_x = Ref{Int64}(-1), Ref{Cdouble}(0.0), Vector{Float64}(undef, post.d)
    GC.@preserve am kinds prm z so sl lb ub best_x begin
        src = isnothing(am.candidates) ? CandSrc(1, 0, C_NULL, am.seed, pointer(lb), pointer(ub), 0) :
                                         CandSrc(0, 0, pointer(am.candidates), 0, C_NULL, C_NULL, 0)
        M = isnothing(am.candidates) ? am.count : size(am.candidates, 2)
        check(post.m, ccall((:bosip_b200_acq_argmax_multi, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ref{CandSrc}, Int64, Int64, Cint, Ref{Int64}, Ref{Cdouble}, Ptr{Cdouble}),
            post.h, Ref(_like_struct(kind, z, so, sl)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)), Ref(src), M, 0, acq,
            best_idx, best_val, best_x))
    end
    return best_x, best_val[]
end
function evidence_multi(post::B200MultiGPPosterior, like::B200Likelihood, x_prior::Product, xs::AbstractMatrix{<:Real}; which::Cuint=WANT_MEAN)
    Xc = Matrix{Float64}(xs); S = size(Xc, 2); out = Ref{Cdouble}(0.0)
    kinds, prm = _prior_desc(x_prior); kind, z, so, sl = _like_desc(like)
    GC.@preserve Xc kinds prm z so sl begin
        check(post.m, ccall((:bosip_b200_evidence_sum_multi, LIB), Cint,
            (Ptr{Cvoid}, Ref{LikeNormal}, Ref{Prior}, Ptr{Cdouble}, Int64, Cuint, Ref{Cdouble}),
            post.h, Ref(_like_struct(kind, z, so, sl)), Ref(Prior(post.d, pointer(kinds), pointer(prm), C_NULL)), Xc, S, which, out))
    end
    return out[] / S
end
function calc_mmd_multi(m::MultiContext, kernel_id::Integer, lengthscales::Vector{Float64}, X::Matrix{Float64}, Y::Matrix{Float64})
    out = Ref{Cdouble}(0.0)
    check(m, ccall((:bosip_b200_mmd_multi, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Ref{Cdouble}),
        m.h, kernel_id, size(X, 1), lengthscales, X, size(X, 2), Y, size(Y, 2), out))
    return out[]
end
function calc_tv_multi(m::MultiContext, log_ws::Vector{Float64}, approx_logvals::Vector{Float64}, true_logvals::Vector{Float64})
    out = Ref{Cdouble}(0.0)
    check(m, ccall((:bosip_b200_tv_multi, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ref{Cdouble}),
        m.h, log_ws, approx_logvals, true_logvals, length(log_ws), out))
    return out[]
end


# ---- variance engine ----------------------------------------------------------------------------------------------
"""
Select the kernel behind the variance half of mean_and_var: :fp64 (DMMA), :int8 (tcgen05 on error-free byte slices) or :auto.
`slices = 0` is the default INT8 scheme (7 bytes of α²L⁻¹ against 6 of the cross-kernel values, 27 exact INT8 GEMMs); 6, 7, 8 select
the symmetric schemes.
"""
function set_variance_engine!(engine::Symbol=:auto; slices::Integer=0)
    ctx = context()
    code = engine === :fp64 ? 0 : engine === :int8 ? 1 : 2
    check(ctx, ccall((:bosip_b200_set_variance_engine, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, code, slices))
end

"The current setting as (engine, slices of α²L⁻¹, slices of K*)."
function variance_engine()
    ctx = context()
    e = Ref{Cint}(0); s = Ref{Cint}(0); a = Ref{Cint}(0)
    check(ctx, ccall((:bosip_b200_get_variance_engine, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), ctx.h, e, s, a))
    return ((:fp64, :int8, :auto)[e[] + 1], Int(s[]), Int(a[]))
end

# ---- AMIS inner loop (src/samplers/amis/amis_method.jl:46-78) -------------------------------------------------------
# A drop-in `fit_distribution!` + weight update for NormalProposal: the sample matrix, log_P, Delta and log_Omega are ordinary Julia
# arrays here (HOST); a maintainer who keeps them in CUDA memory passes DEVICE pointers instead.
function BOSIP.estimate_parameters!(dist::BOSIP.NormalProposal, xs::Matrix{Float64}, log_ws::Vector{Float64}, ::Val{:b200})
    ctx = context(); d = size(xs, 1)
    mean = zeros(d); cov = zeros(d, d); sums = zeros(2)
    check(ctx, ccall((:bosip_b200_weighted_moments, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        ctx.h, d, xs, log_ws, size(xs, 2), HOST, mean, cov, sums, C_NULL))
    Σ = Symmetric(cov)
    (any(isinf.(Σ)) || !isposdef(Σ)) && return dist.dist          # normal.jl:92-95: keep the previous proposal
    dist.dist = MvNormal(mean, Σ)
end
"Δ .+= Σ_q pdf.(Ref(q), eachcol(xs)) for a vector of MvNormal proposals (amis_method.jl:66-68,73-74)."
function mixture_pdf_sum!(Δ::Vector{Float64}, qs::Vector{<:MvNormal}, xs::Matrix{Float64})
    ctx = context(); d = size(xs, 1); nq = length(qs)
    mus = hcat(mean.(qs)...)
    Ls = [cholesky(Symmetric(Matrix(cov(q)))).L for q in qs]
    Linvs = cat((Matrix(inv(L)) for L in Ls)...; dims=3)
    lognorm = [-sum(log.(diag(L))) - d / 2 * log(2π) for L in Ls]
    check(ctx, ccall((:bosip_b200_mixture_pdf_sum, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Cint, Ptr{Cdouble}),
        ctx.h, d, nq, mus, Linvs, lognorm, xs, size(xs, 2), HOST, 1, Δ))
    return Δ
end
"resample(xs, ws, count) (src/sampling.jl:55-57)"
function resample_b200(xs::Matrix{Float64}, ws::Vector{Float64}, count::Integer; seed::Integer=0)
    ctx = context(); out = Matrix{Float64}(undef, size(xs, 1), count)
    check(ctx, ccall((:bosip_b200_resample, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64, UInt64, Int64, Cint, Ptr{Cdouble}, Ptr{Int64}),
        ctx.h, size(xs, 1), xs, ws, size(xs, 2), seed, count, HOST, out, C_NULL))
    return out
end

end # module
