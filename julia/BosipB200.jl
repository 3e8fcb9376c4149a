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


# ---- variance engine ----------------------------------------------------------------------------------------------
"Select the kernel behind the variance half of mean_and_var: :fp64 (DMMA), :int8 (tcgen05 on error-free byte slices) or :auto."
function set_variance_engine!(engine::Symbol=:auto; slices::Integer=0)
    ctx = context()
    code = engine === :fp64 ? 0 : engine === :int8 ? 1 : 2
    check(ctx, ccall((:bosip_b200_set_variance_engine, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, code, slices))
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
