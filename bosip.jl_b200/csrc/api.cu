// C ABI of libbosip_b200.so (include/bosip_b200.h): context lifecycle, the chunked evaluation pipeline
// (H2D staging | cross-kernel+mean | variance GEMM | fused epilogue | reductions | D2H), and the entry points.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace bosip {

static thread_local std::string g_init_error;

// ------------------------------------------------------------------------------------------------ timing events
struct EvPair { cudaEvent_t a, b; int kind; };
struct TimingPool {
  std::vector<EvPair> pool;
  size_t used = 0;
};
static TimingPool* pool_of(Ctx* ctx);

}  // namespace bosip

struct bosip_b200_ctx {
  bosip::Ctx c;
  bosip::TimingPool tp;
};
struct bosip_b200_gp {
  bosip::Gp* g;
};

namespace bosip {

static TimingPool* pool_of(Ctx* ctx) { return &reinterpret_cast<bosip_b200_ctx*>(ctx)->tp; }

static void t_begin(Ctx* ctx, int kind, cudaStream_t st) {
  if (!ctx->timing.enabled) return;
  TimingPool* tp = pool_of(ctx);
  if (tp->used == tp->pool.size()) {
    EvPair e; cudaEventCreate(&e.a); cudaEventCreate(&e.b); e.kind = kind;
    tp->pool.push_back(e);
  }
  tp->pool[tp->used].kind = kind;
  cudaEventRecord(tp->pool[tp->used].a, st);
}
static void t_end(Ctx* ctx, cudaStream_t st) {
  if (!ctx->timing.enabled) return;
  TimingPool* tp = pool_of(ctx);
  cudaEventRecord(tp->pool[tp->used].b, st);
  tp->used++;
}
static void t_collect(Ctx* ctx) {  // call after the stream has been synchronised
  if (!ctx->timing.enabled) return;
  TimingPool* tp = pool_of(ctx);
  for (size_t i = 0; i < tp->used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, tp->pool[i].a, tp->pool[i].b) == cudaSuccess) {
      ctx->timing.ms[tp->pool[i].kind] += ms;
      ctx->timing.launches[tp->pool[i].kind] += 1;
    }
  }
  tp->used = 0;
}

// ------------------------------------------------------------------------------------------------ helpers
int ensure_ws(Ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return 0;
  if (ctx->ws) { cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_bytes = 0; }
  BOSIP_CUDA(ctx, cudaMalloc(&ctx->ws, bytes));
  ctx->ws_bytes = bytes;
  return 0;
}

static int ensure_pin(Ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pin_bytes) return 0;
  if (ctx->pin) { cudaFreeHost(ctx->pin); ctx->pin = nullptr; ctx->pin_bytes = 0; }
  BOSIP_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->pin), bytes, cudaHostAllocDefault));
  ctx->pin_bytes = bytes;
  return 0;
}
// true for plain malloc'ed host memory (not cudaHostAlloc'ed / cudaHostRegister'ed): copies from/to it are staged by the driver
// and BLOCK the calling thread, a device-to-host copy until the producing kernels have finished
static bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

int make_like_params(Ctx* ctx, const bosip_b200_like_normal* like, int gp_p, EpiParams* ep) {
  const int q = like->p, kind = like->kind;
  if (q < 1 || q > MAX_P) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood dimension out of range");
  ep->like_kind = kind; ep->q = q; ep->c_approx = ep->c_mean = ep->c_sq = 0.0; ep->log_C = 0.0;
  for (int i = 0; i < q; ++i) { ep->grp[i] = 1; ep->z[i] = 0.0; ep->so[i] = 1.0; }
  if (kind == BOSIP_B200_LIKE_EXP) {  // src/likelihoods/exp.jl:9-16: exactly one model output
    if (q != 1 || gp_p != 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "ExpLikelihood needs a single model output");
    return 0;
  }
  if (kind < 0 || kind > BOSIP_B200_LIKE_EXP) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown likelihood kind");
  if (!like->std_obs) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood needs std_obs");
  if (kind != BOSIP_B200_LIKE_NORMAL_DIFF && !like->z_obs) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood needs z_obs");
  int total = q;
  if (kind == BOSIP_B200_LIKE_NORMAL_SUM) {
    if (!like->sum_lengths) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "NormalSumLikelihood needs sum_lengths");
    total = 0;
    for (int i = 0; i < q; ++i) {
      if (like->sum_lengths[i] < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "sum_lengths must be >= 1");
      ep->grp[i] = like->sum_lengths[i]; total += like->sum_lengths[i];
    }
  }
  if (total != gp_p) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood dimension p does not match the GP's output dimension");
  double logc = 0.0, sum_logz = 0.0;
  for (int i = 0; i < q; ++i) {
    double so = like->std_obs[i], z = (kind == BOSIP_B200_LIKE_NORMAL_DIFF) ? 0.0 : like->z_obs[i];
    if (kind == BOSIP_B200_LIKE_LOGNORMAL) {
      // sigma_log = sqrt(log(1 + CV^2)); mu_log = log_y - sigma_log^2/2 (lognormal.jl:44-45).  logpdf(LogNormal(mu_log, s), z)
      // = logN(log z; log_y - sigma_log^2/2, s) - log z = logN(log z + sigma_log^2/2; log_y, s) - log z
      if (!(so > 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "CV must be > 0");
      so = sqrt(log(1.0 + so * so));
      sum_logz += z;
      z = z + so * so / 2.0;
    }
    if (!(so > 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "std_obs must be > 0");
    ep->z[i] = z; ep->so[i] = so;
    logc += log(2.0 * sqrt(M_PI) * so);  // src/likelihoods/normal.jl:75
  }
  ep->log_C = -logc;
  if (kind == BOSIP_B200_LIKE_LOGNORMAL) { ep->c_approx = ep->c_mean = -sum_logz; ep->c_sq = -2.0 * sum_logz; }  // lognormal.jl:47-50,84
  return 0;
}

static int make_epi(Ctx* ctx, const Gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior, EpiParams* ep) {
  memset(ep, 0, sizeof(*ep));
  ep->p = gp->p; ep->d = gp->d;
  ep->pred_noise = gp->opts.pred_noise; ep->clamp_var = gp->opts.clamp_var;
  for (int j = 0; j < gp->p; ++j) { ep->kxx[j] = gp->alpha2[j]; ep->sig2[j] = gp->sig2[j]; ep->z[j] = 0.0; ep->so[j] = 1.0; ep->grp[j] = 1; }
  ep->q = gp->p;
  if (like) BOSIP_TRY(make_like_params(ctx, like, gp->p, ep));
  ep->has_prior = 0;
  if (prior) {
    if (prior->logprior) ep->has_prior = 2;
    else if (prior->kind) {
      if (prior->d != gp->d) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "prior dimension d does not match the GP's input dimension");
      if (!prior->params) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "prior descriptor needs params");
      ep->has_prior = 1;
      for (int k = 0; k < gp->d; ++k) {
        const double* q = prior->params + 4 * k;
        ep->prior_kind[k] = prior->kind[k];
        for (int c = 0; c < 4; ++c) ep->prior_prm[k][c] = q[c];
        switch (prior->kind[k]) {
          case BOSIP_B200_PRIOR_UNIFORM:
            if (!(q[1] > q[0])) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "uniform prior needs a < b");
            ep->prior_const[k] = -log(q[1] - q[0]);
            break;
          case BOSIP_B200_PRIOR_NORMAL:
            if (!(q[1] > 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "normal prior needs s > 0");
            break;
          case BOSIP_B200_PRIOR_TRUNCNORMAL: {
            if (!(q[1] > 0.0) || !(q[3] > q[2])) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "truncated normal prior needs s > 0 and a < b");
            // Distributions.Truncated: logtp = log(Phi((b-m)/s) - Phi((a-m)/s))
            const double tp = 0.5 * erfc(-(q[3] - q[0]) / q[1] * M_SQRT1_2) - 0.5 * erfc(-(q[2] - q[0]) / q[1] * M_SQRT1_2);
            ep->prior_const[k] = -log(tp);
            break;
          }
          default: BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown prior kind");
        }
      }
    }
  }
  return 0;
}

// What one evaluation sweep does.  Pointers in "user space" follow `mem`.
struct Sweep {
  Gp* gp = nullptr;
  std::vector<Gp*> fits;          // BI ensemble (hyper-parameter samples); empty => {gp}
  std::vector<EpiParams> fit_ep;   // per-sample epilogue parameters (kxx, sigma^2, options differ per sample)
  EpiParams ep;
  bool use_like = false;
  int mem = BOSIP_B200_HOST;
  const double* xq = nullptr;                 // d x M (user space) or nullptr with src
  const bosip_b200_cand_src* src = nullptr;   // generated candidates
  int64_t M = 0, index_offset = 0;
  const double* mean_at_xq = nullptr;         // p x M user space
  const double* logprior = nullptr;           // M user space
  bool need_var = false;
  double *mu = nullptr, *var = nullptr, *la = nullptr, *lm = nullptr, *lv = nullptr;  // user-space outputs
  bool do_argmax = false; int exp_vals = 0;
  bool do_evidence = false; unsigned evid_which = 0;
  int64_t seg_len = 0; double* seg_out = nullptr; unsigned seg_which = 0;  // segmented mean of exp(closure): seg_out (device) [M / seg_len]
  // results
  int64_t n_clamped = 0; int64_t domain_idx = -1;
  double best_val = 0.0; int64_t best_idx = -1; double evid_sum = 0.0;
};

static int run_sweep(Ctx* ctx, Sweep& sw) {
  if (sw.fits.empty()) { sw.fits.push_back(sw.gp); sw.fit_ep.push_back(sw.ep); }
  Gp* gp = sw.fits[0];
  const int n_fits = (int)sw.fits.size();
  const int p = gp->p, d = gp->d, Ns = gp->Ns;
  const int64_t M = sw.M;
  if (M < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "M must be >= 0");
  if (M == 0) return 0;
  if (!sw.xq && !sw.src) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "query points are required");
  const bool host = (sw.mem == BOSIP_B200_HOST);
  const bool gen = (sw.xq == nullptr);

  // ---- variance engine (FP64 DMMA or INT8 tcgen05 with error-free slicing)
  // AUTO: the INT8 engine pays off once the contraction is more than a tile deep; below that the FP64 kernel is as fast
  bool i8 = sw.need_var && (ctx->var_engine == BOSIP_B200_VAR_INT8 || (ctx->var_engine == BOSIP_B200_VAR_AUTO && gp->N >= 128));
  if (i8 && ctx->var_engine == BOSIP_B200_VAR_AUTO && ((double)gp->N * ctx->var_slices * 255.0 * 128.0 >= 2147483647.0 || gp->N >= 8192)) i8 = false;
  const int S8 = ctx->var_slices, SA8 = ctx->var_a_slices, KB8 = (gp->N + 127) / 128;   // slices of alpha^2 L^-1 / of K*
  int nsub8 = 1;
  if (i8) {
    for (Gp* gf : sw.fits) BOSIP_TRY(i8_prepare(ctx, gf, S8, ctx->s_comp));
    nsub8 = i8_nsub(gp);
  }

  // ---- n-split of the cross-kernel so that its shared-memory footprint allows >= 4 CTAs per SM
  // (the INT8 slice writer walks 32 training points per iteration: spans are multiples of 32)
  static const int kstar_smem_kb = getenv("BOSIP_KSTAR_SMEM_KB") ? atoi(getenv("BOSIP_KSTAR_SMEM_KB")) : 40;  // dev A/B switch; measured per 8 / 4 / 14 launches
  // at C3 / C5 / C2 (ms): 16 KB 6.21 / 4.17 / 1.27, 24 KB 6.28 / 4.33 / 1.29, 32 KB 6.09 / 4.41 / 1.19, 40 KB 5.93 / 4.48 / 1.09, 48 KB 6.38 / 4.53 / 1.10, 56 KB 6.42 / 4.45 / 1.24
  const int span_q = 32;  // same partition for every engine and for the mean-only path: the mean stays bit-identical across them
  int nspan = (int)((kstar_smem_kb * 1024) / (8 * (gp->dp + 1))) & ~(span_q - 1);
  nspan = std::max(span_q, std::min(nspan, (int)round_up(Ns, span_q)));
  int nsplit = (Ns + nspan - 1) / nspan;
  nspan = (int)round_up((Ns + nsplit - 1) / nsplit, span_q);
  nsplit = (Ns + nspan - 1) / nspan;

  // ---- chunk size
  int64_t mc;
  if (sw.need_var) {
    const double k_bytes_per_point = i8 ? (double)p * KB8 * 128.0 * SA8 : (double)p * Ns * 8.0;
    static const double ws_cap = getenv("BOSIP_CHUNK_WS_GB") ? atof(getenv("BOSIP_CHUNK_WS_GB")) * 1e9 : 3.0e9;   // dev A/B switch
    const int64_t by_mem = (int64_t)(ws_cap / k_bytes_per_point) / TILE_M * TILE_M;
    static const int tiles_per_sm = getenv("BOSIP_CHUNK_TILES_PER_SM") ? std::max(1, atoi(getenv("BOSIP_CHUNK_TILES_PER_SM"))) : 4;   // dev A/B switch
    mc = std::min<int64_t>((int64_t)ctx->sms * TILE_M * tiles_per_sm, std::max<int64_t>(TILE_M, by_mem));
  } else {
    mc = (int64_t)ctx->sms * TILE_M * 16;
  }
  if (ctx->user_chunk > 0) mc = std::min(mc, round_up(ctx->user_chunk, TILE_M));
  mc = std::min(mc, round_up(M, TILE_M));
  if (sw.seg_len > mc) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "segment length exceeds the chunk size");
  const int64_t ppc = sw.seg_len > 0 ? (mc / sw.seg_len) * sw.seg_len : mc;  // points per chunk: whole segments only
  const int64_t nchunks = (M + ppc - 1) / ppc;

  // ---- workspace carve-up (all sizes in doubles, 256-byte aligned pieces)
  size_t off = 0;
  auto carve = [&](size_t n_doubles) { size_t o = off; off += (n_doubles * 8 + 255) / 256 * 256; return o; };
  const size_t o_Ks = sw.need_var ? carve(i8 ? (size_t)p * (mc / TILE_M) * KB8 * SA8 * 2048 : (size_t)p * Ns * mc) : 0;
  const size_t o_mu_part = carve((size_t)nsplit * p * mc);
  const size_t o_q_part = sw.need_var ? carve((size_t)p * mc) : 0;
  const size_t o_q_scr = i8 ? carve((size_t)nsub8 * p * mc) : 0;
  const bool stage_x = host || gen;
  size_t o_x[2] = {0, 0}, o_mean[2] = {0, 0}, o_lpin[2] = {0, 0}, o_mu[2] = {0, 0}, o_var[2] = {0, 0}, o_la[2] = {0, 0},
         o_lm[2] = {0, 0}, o_lv[2] = {0, 0}, o_lp[2] = {0, 0};
  const bool want_lv_dev = sw.lv || sw.do_argmax;
  const bool want_post_dev = sw.do_evidence;
  for (int b = 0; b < 2; ++b) {
    if (stage_x) o_x[b] = carve((size_t)d * mc);
    if (host && sw.mean_at_xq) o_mean[b] = carve((size_t)p * mc);
    if (host && sw.logprior) o_lpin[b] = carve((size_t)mc);
    if (host && sw.mu) o_mu[b] = carve((size_t)p * mc);
    if (host && sw.var) o_var[b] = carve((size_t)p * mc);
    if ((host && sw.la) || (want_post_dev && sw.evid_which == BOSIP_B200_WANT_APPROX) || (sw.seg_len > 0 && sw.seg_which == BOSIP_B200_WANT_APPROX)) o_la[b] = carve((size_t)mc);
    if ((host && sw.lm) || (want_post_dev && sw.evid_which == BOSIP_B200_WANT_MEAN) || (sw.seg_len > 0 && sw.seg_which == BOSIP_B200_WANT_MEAN)) o_lm[b] = carve((size_t)mc);
    if ((host && sw.lv) || (sw.do_argmax && !sw.lv) || (sw.seg_len > 0 && sw.seg_which == BOSIP_B200_WANT_VAR)) o_lv[b] = carve((size_t)mc);
    if (want_post_dev) o_lp[b] = carve((size_t)mc);
  }
  const size_t o_scr_v = carve(1024), o_scr_i = carve(1024), o_scr_s = carve(256), o_scalars = carve(8);
  BOSIP_TRY(ensure_ws(ctx, off));
  char* ws = ctx->ws;
  auto D = [&](size_t o) { return reinterpret_cast<double*>(ws + o); };
  double* Ks = sw.need_var ? D(o_Ks) : nullptr;
  double* mu_part = D(o_mu_part);
  double* q_part = sw.need_var ? D(o_q_part) : nullptr;
  // scalars: [0] best_val, [1] best_idx (as long long), [2] evidence accum, [3] n_clamped (ull), [4] domain idx (ull)
  double* scal = D(o_scalars);
  {
    unsigned long long init[8];
    double zero = 0.0; long long minus1 = -1;
    memcpy(&init[0], &zero, 8); memcpy(&init[1], &minus1, 8); memcpy(&init[2], &zero, 8);
    init[3] = 0ull; init[4] = ~0ull; init[5] = init[6] = init[7] = 0ull;
    BOSIP_CUDA(ctx, cudaMemcpyAsync(scal, init, sizeof(init), cudaMemcpyHostToDevice, ctx->s_comp));
  }
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(scal + 3);

  cudaStream_t sc = ctx->s_comp, si = ctx->s_in, so = ctx->s_out;
  (void)want_lv_dev;

  // ---- pageable host buffers (what a Julia `Array{Float64}` is): a cudaMemcpyAsync to or from them blocks the calling thread -- a
  // device-to-host copy until the chunk's kernels have finished -- so the launch loop could never run ahead of the GPU (measured:
  // -6 % end to end).  They are staged through library-owned pinned double buffers instead: host memcpy user -> pinned before the
  // H2D of chunk c, pinned -> user for chunk c-2 once its D2H event has fired; still ONE cudaMemcpyAsync per array per chunk.
  const bool use_out = sw.use_like;
  const bool pin_stage = host && (is_pageable(sw.xq) || is_pageable(sw.mean_at_xq) || is_pageable(sw.logprior) || is_pageable(sw.mu) ||
                                  is_pageable(sw.var) || (use_out && (is_pageable(sw.la) || is_pageable(sw.lm) || is_pageable(sw.lv))));
  size_t pi_x[2] = {0, 0}, pi_mean[2] = {0, 0}, pi_lp[2] = {0, 0}, po_mu[2] = {0, 0}, po_var[2] = {0, 0}, po_la[2] = {0, 0}, po_lm[2] = {0, 0},
         po_lv[2] = {0, 0};
  if (pin_stage) {
    size_t poff = 0;
    auto pcarve = [&](size_t n_doubles) { size_t o = poff; poff += (n_doubles * 8 + 255) / 256 * 256; return o; };
    for (int b = 0; b < 2; ++b) {
      pi_x[b] = pcarve((size_t)d * mc);
      if (sw.mean_at_xq) pi_mean[b] = pcarve((size_t)p * mc);
      if (sw.logprior) pi_lp[b] = pcarve((size_t)mc);
      if (sw.mu) po_mu[b] = pcarve((size_t)p * mc);
      if (sw.var) po_var[b] = pcarve((size_t)p * mc);
      if (sw.la && use_out) po_la[b] = pcarve((size_t)mc);
      if (sw.lm && use_out) po_lm[b] = pcarve((size_t)mc);
      if (sw.lv && use_out) po_lv[b] = pcarve((size_t)mc);
    }
    BOSIP_TRY(ensure_pin(ctx, poff));
  }
  auto PIN = [&](size_t o) { return reinterpret_cast<double*>(ctx->pin + o); };
  // copy chunk `cc`'s outputs from pinned buffer bb to the caller's arrays (its D2H event must have completed)
  auto drain_out = [&](int64_t cc, int bb) {
    const int64_t q0 = cc * ppc, qv = std::min(ppc, M - q0);
    if (sw.mu) memcpy(sw.mu + (size_t)q0 * p, PIN(po_mu[bb]), (size_t)qv * p * 8);
    if (sw.var) memcpy(sw.var + (size_t)q0 * p, PIN(po_var[bb]), (size_t)qv * p * 8);
    if (sw.la && use_out) memcpy(sw.la + q0, PIN(po_la[bb]), (size_t)qv * 8);
    if (sw.lm && use_out) memcpy(sw.lm + q0, PIN(po_lm[bb]), (size_t)qv * 8);
    if (sw.lv && use_out) memcpy(sw.lv + q0, PIN(po_lv[bb]), (size_t)qv * 8);
  };

  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t c0 = c * ppc, mv = std::min(ppc, M - c0);
    const int64_t m_tiles = (mv + TILE_M - 1) / TILE_M, mcu = m_tiles * TILE_M;  // rows actually computed this chunk
    const double *x_dev, *mean_dev = nullptr, *lpin_dev = nullptr;
    // ---------------- inputs
    if (gen) {
      if (c >= 2) BOSIP_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_out[b], 0));
      BOSIP_TRY(launch_candidates(ctx, sw.src, d, mv, sw.index_offset + c0, D(o_x[b]), sc));
      x_dev = D(o_x[b]);
    } else if (host) {
      const double *hx = sw.xq + (size_t)c0 * d, *hmean = sw.mean_at_xq ? sw.mean_at_xq + (size_t)c0 * p : nullptr,
                   *hlp = sw.logprior ? sw.logprior + c0 : nullptr;
      if (pin_stage) {
        if (c >= 2) {   // chunk c-2 used the same pinned buffers: its copies are done once its D2H event has fired
          BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_out[b]));
          drain_out(c - 2, b);
        }
        memcpy(PIN(pi_x[b]), hx, (size_t)mv * d * 8); hx = PIN(pi_x[b]);
        if (hmean) { memcpy(PIN(pi_mean[b]), hmean, (size_t)mv * p * 8); hmean = PIN(pi_mean[b]); }
        if (hlp) { memcpy(PIN(pi_lp[b]), hlp, (size_t)mv * 8); hlp = PIN(pi_lp[b]); }
      }
      if (c >= 2) BOSIP_CUDA(ctx, cudaStreamWaitEvent(si, ctx->ev_comp[b], 0));  // buffer b's previous consumer is done
      BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_x[b]), hx, (size_t)mv * d * 8, cudaMemcpyHostToDevice, si));
      if (hmean) BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_mean[b]), hmean, (size_t)mv * p * 8, cudaMemcpyHostToDevice, si));
      if (hlp) BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_lpin[b]), hlp, (size_t)mv * 8, cudaMemcpyHostToDevice, si));
      BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_in[b], si));
      BOSIP_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_in[b], 0));
      if (c >= 2) BOSIP_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_out[b], 0));  // output buffer b has been drained
      x_dev = D(o_x[b]);
      if (sw.mean_at_xq) mean_dev = D(o_mean[b]);
      if (sw.logprior) lpin_dev = D(o_lpin[b]);
    } else {
      x_dev = sw.xq + (size_t)c0 * d;
      if (sw.mean_at_xq) mean_dev = sw.mean_at_xq + (size_t)c0 * p;
      if (sw.logprior) lpin_dev = sw.logprior + c0;
    }
    // ---------------- compute
    ctx->launches += n_fits * (2 + (sw.need_var ? 1 : 0)) + (n_fits > 1 ? 1 : 0) + (gen ? 1 : 0) + (sw.do_argmax ? 2 : 0) + (sw.do_evidence ? 2 : 0);
    double* mu_o = sw.mu ? (host ? D(o_mu[b]) : sw.mu + (size_t)c0 * p) : nullptr;
    double* var_o = sw.var ? (host ? D(o_var[b]) : sw.var + (size_t)c0 * p) : nullptr;
    double* la_o = o_la[b] ? D(o_la[b]) : (sw.la ? sw.la + c0 : nullptr);
    double* lm_o = o_lm[b] ? D(o_lm[b]) : (sw.lm ? sw.lm + c0 : nullptr);
    double* lv_o = o_lv[b] ? D(o_lv[b]) : (sw.lv ? sw.lv + c0 : nullptr);
    double* lp_o = o_lp[b] ? D(o_lp[b]) : nullptr;
    if (!sw.use_like) { la_o = lm_o = lv_o = lp_o = nullptr; }
    for (int f = 0; f < n_fits; ++f) {
      Gp* gf = sw.fits[f];
      t_begin(ctx, 0, sc);
      BOSIP_TRY(launch_kstar(ctx, gf, x_dev, mv, mc, m_tiles, nsplit, nspan, Ks, mu_part, sw.need_var ? (i8 ? 2 : 1) : 0, SA8, sc));
      t_end(ctx, sc);
      if (sw.need_var) {
#ifdef BOSIP_I8_DEBUG_FILL_A   // dev build: overwrite the K* slices with a constant byte before the GEMM (energy-per-operation probe; results are garbage)
        if (i8) { static const int fill = getenv("BOSIP_I8_FILL_A") ? atoi(getenv("BOSIP_I8_FILL_A")) : -1;
                  if (fill >= 0) cudaMemsetAsync(Ks, fill, (size_t)p * (mc / TILE_M) * KB8 * SA8 * 16384, sc); }
#endif
        t_begin(ctx, 1, sc);
        if (i8) {
          BOSIP_TRY(launch_vargemm_i8(ctx, gf, reinterpret_cast<const uint8_t*>(Ks), SA8, mc, m_tiles, D(o_q_scr), q_part, sc));
          ctx->launches += 1;
        } else {
          BOSIP_TRY(launch_vargemm(ctx, gf, Ks, mc, m_tiles, q_part, sc));
        }
        t_end(ctx, sc);
      }
      t_begin(ctx, 2, sc);
      BOSIP_TRY(launch_epilogue(ctx, gf, sw.fit_ep[f], x_dev, lpin_dev, mean_dev, mv, mc, nsplit, mu_part, sw.need_var, q_part, mu_o, var_o,
                                lp_o, la_o, lm_o, lv_o, counters, (uint64_t)c0, n_fits == 1 ? 0 : (f == 0 ? 1 : 2), sc));
      if (f + 1 < n_fits) t_end(ctx, sc);
    }
    if (n_fits > 1) BOSIP_TRY(launch_bi_finish(ctx, sw.ep, x_dev, lpin_dev, mv, n_fits, lp_o, la_o, lm_o, lv_o, sc));
    (void)mcu;
    if (sw.do_argmax)
      BOSIP_TRY(launch_argmax_chunk(ctx, lv_o, mv, sw.index_offset + c0, sw.exp_vals, D(o_scr_v), reinterpret_cast<long long*>(D(o_scr_i)),
                                    scal + 0, reinterpret_cast<long long*>(scal + 1), sc));
    if (sw.seg_len > 0) {
      const double* v = sw.seg_which == BOSIP_B200_WANT_APPROX ? la_o : sw.seg_which == BOSIP_B200_WANT_MEAN ? lm_o : lv_o;
      BOSIP_TRY(launch_segmean_exp(ctx, v, mv / sw.seg_len, sw.seg_len, sw.seg_out + c0 / sw.seg_len, sc));
      ctx->launches += 1;
    }
    if (sw.do_evidence)
      BOSIP_TRY(launch_sum_chunk(ctx, sw.evid_which == BOSIP_B200_WANT_APPROX ? la_o : lm_o, lp_o, mv, D(o_scr_s), scal + 2, sc));
    t_end(ctx, sc);
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_comp[b], sc));
    // ---------------- outputs
    if (host) {
      BOSIP_CUDA(ctx, cudaStreamWaitEvent(so, ctx->ev_comp[b], 0));
      if (sw.mu) BOSIP_CUDA(ctx, cudaMemcpyAsync(pin_stage ? PIN(po_mu[b]) : sw.mu + (size_t)c0 * p, D(o_mu[b]), (size_t)mv * p * 8, cudaMemcpyDeviceToHost, so));
      if (sw.var) BOSIP_CUDA(ctx, cudaMemcpyAsync(pin_stage ? PIN(po_var[b]) : sw.var + (size_t)c0 * p, D(o_var[b]), (size_t)mv * p * 8, cudaMemcpyDeviceToHost, so));
      if (sw.la && sw.use_like) BOSIP_CUDA(ctx, cudaMemcpyAsync(pin_stage ? PIN(po_la[b]) : sw.la + c0, D(o_la[b]), (size_t)mv * 8, cudaMemcpyDeviceToHost, so));
      if (sw.lm && sw.use_like) BOSIP_CUDA(ctx, cudaMemcpyAsync(pin_stage ? PIN(po_lm[b]) : sw.lm + c0, D(o_lm[b]), (size_t)mv * 8, cudaMemcpyDeviceToHost, so));
      if (sw.lv && sw.use_like) BOSIP_CUDA(ctx, cudaMemcpyAsync(pin_stage ? PIN(po_lv[b]) : sw.lv + c0, D(o_lv[b]), (size_t)mv * 8, cudaMemcpyDeviceToHost, so));
    }
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_out[b], so));
  }
  if (pin_stage) {   // the last two chunks still sit in the pinned buffers
    for (int64_t c = std::max<int64_t>(0, nchunks - 2); c < nchunks; ++c) {
      BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_out[c & 1]));
      drain_out(c, (int)(c & 1));
    }
  }
  unsigned long long res[8];
  BOSIP_CUDA(ctx, cudaMemcpyAsync(res, scal, sizeof(res), cudaMemcpyDeviceToHost, sc));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(sc));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(so));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(si));
  t_collect(ctx);
  memcpy(&sw.best_val, &res[0], 8);
  long long bi; memcpy(&bi, &res[1], 8); sw.best_idx = bi;
  memcpy(&sw.evid_sum, &res[2], 8);
  sw.n_clamped = (int64_t)res[3];
  sw.domain_idx = (res[4] == ~0ull) ? -1 : (int64_t)res[4];
  if (sw.domain_idx >= 0)
    BOSIP_FAIL(ctx, BOSIP_B200_EDOMAIN,
               "DomainError: likelihood variance below -1e-8 (src/posterior/likelihood.jl:61-65) at query index " +
                   std::to_string(sw.domain_idx));
  return 0;
}

// Fill sw.gp / sw.ep (one fit) or sw.fits / sw.fit_ep (BI ensemble of hyper-parameter samples).
static int set_fits(Ctx* ctx, Sweep& sw, bosip_b200_gp* const* gps, int n, const bosip_b200_like_normal* like, const bosip_b200_prior* prior) {
  sw.gp = gps[0]->g;
  if (n == 1) return make_epi(ctx, sw.gp, like, prior, &sw.ep);
  for (int f = 0; f < n; ++f) {
    if (!gps[f] || !gps[f]->g) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "NULL GP handle in the ensemble");
    Gp* g = gps[f]->g;
    if (g->ctx != sw.gp->ctx || g->d != sw.gp->d || g->p != sw.gp->p || g->N != sw.gp->N || g->kernel_id != sw.gp->kernel_id)
      BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "all hyper-parameter samples must share context, kernel, d, p and N");
    EpiParams e;
    BOSIP_TRY(make_epi(ctx, g, like, prior, &e));
    sw.fits.push_back(g); sw.fit_ep.push_back(e);
  }
  sw.ep = sw.fit_ep[0];  // carries the prior for bi_finish
  return 0;
}

}  // namespace bosip

using namespace bosip;

// Makes the context's device current for the duration of an entry point and restores the caller's device afterwards: a host that
// drives several contexts from one thread (bosip_b200_init_multi, or a framework with its own notion of "current device" such as
// torch) must not find its device changed behind its back.
struct DeviceScope {
  int prev = -1;
  bool ok;
  explicit DeviceScope(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; ok = cudaSetDevice(dev) == cudaSuccess; }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

#define API_LOCK(ctxp)                        \
  if (!(ctxp)) return BOSIP_B200_EINVAL;      \
  Ctx* ctx = &(ctxp)->c;                      \
  std::lock_guard<std::mutex> lock__(ctx->mu); \
  DeviceScope dev_scope__(ctx->device);       \
  if (!dev_scope__.ok) { ctx->err = "cudaSetDevice failed"; return BOSIP_B200_ECUDA; }

extern "C" {

int bosip_b200_version(void) { return BOSIP_B200_VERSION; }

const char* bosip_b200_last_error(const bosip_b200_ctx* c) { return c ? c->c.err.c_str() : g_init_error.c_str(); }

int bosip_b200_init(int device, bosip_b200_ctx** out) {
  if (!out) return BOSIP_B200_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_init_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libbosip_b200 has no CPU path";
    return BOSIP_B200_ECUDA;
  }
  if (device < 0 || device >= ndev) { g_init_error = "device index out of range"; return BOSIP_B200_EINVAL; }
  cudaDeviceProp prop;
  DeviceScope dev_scope(device);   // the caller's current device is restored on every return path
  if (!dev_scope.ok || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    if (!dev_scope.ok) e = cudaGetLastError();
    g_init_error = std::string("cudaSetDevice/GetDeviceProperties: ") + cudaGetErrorString(e);
    return BOSIP_B200_ECUDA;
  }
  if (prop.major != 10) {
    g_init_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   "; this library contains sm_100a code only";
    return BOSIP_B200_ECUDA;
  }
  bosip_b200_ctx* h = new bosip_b200_ctx();
  Ctx* ctx = &h->c;
  ctx->device = device; ctx->sms = prop.multiProcessorCount;
  ctx->var_engine = BOSIP_B200_VAR_AUTO;
  if (const char* e = getenv("BOSIP_B200_VAR_ENGINE")) {  // process-wide default, e.g. to run a whole test suite on one engine
    if (!strcmp(e, "fp64")) ctx->var_engine = BOSIP_B200_VAR_FP64;
    else if (!strcmp(e, "int8")) ctx->var_engine = BOSIP_B200_VAR_INT8;
  }
  bool ok = cudaStreamCreateWithFlags(&ctx->s_own, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess;
  ctx->s_comp = ctx->s_own;
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_comp[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->ev_t0) == cudaSuccess && cudaEventCreate(&ctx->ev_t1) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->ev_dep, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { g_init_error = "failed to create CUDA streams/events"; delete h; return BOSIP_B200_ECUDA; }
  *out = h;
  return 0;
}

int bosip_b200_shutdown(bosip_b200_ctx* h) {
  if (!h) return BOSIP_B200_EINVAL;
  Ctx* ctx = &h->c;
  DeviceScope dev_scope(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  for (auto& e : h->tp.pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (int b = 0; b < 2; ++b) { cudaEventDestroy(ctx->ev_in[b]); cudaEventDestroy(ctx->ev_comp[b]); cudaEventDestroy(ctx->ev_out[b]); }
  cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1); cudaEventDestroy(ctx->ev_dep);
  cudaStreamDestroy(ctx->s_own); cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out);
  delete h;
  return 0;
}

int bosip_b200_set_stream(bosip_b200_ctx* h, void* stream) {
  API_LOCK(h);
  BOSIP_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
  ctx->s_comp = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->s_own;
  return 0;
}

int bosip_b200_stream_wait(bosip_b200_ctx* h, void* producer) {
  API_LOCK(h);
  cudaStream_t ps = reinterpret_cast<cudaStream_t>(producer);
  if (ps == ctx->s_comp) return 0;   // same stream: already ordered
  BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_dep, ps));
  BOSIP_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comp, ctx->ev_dep, 0));
  return 0;
}

int bosip_b200_i8_issuers(bosip_b200_ctx* h, int* out) {
  API_LOCK(h);
  if (!out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null output");
  *out = ctx->i8_issuers;
  return 0;
}

int bosip_b200_i8_pair(bosip_b200_ctx* h, int* out) {
  API_LOCK(h);
  if (!out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null output");
  *out = ctx->i8_pair;
  return 0;
}

int bosip_b200_launch_count(bosip_b200_ctx* h, int64_t* out) {
  API_LOCK(h);
  if (!out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null output");
  *out = ctx->launches;
  return 0;
}

int bosip_b200_set_variance_engine(bosip_b200_ctx* h, int engine, int slices) {
  API_LOCK(h);
  if (engine != BOSIP_B200_VAR_FP64 && engine != BOSIP_B200_VAR_INT8 && engine != BOSIP_B200_VAR_AUTO)
    BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown variance engine");
  int a_slices = slices;
  if (slices == 0) { slices = 7; a_slices = 6; }   // default: 7 slices of alpha^2 L^-1 against 6 of K* (27 slice pairs)
  if (slices < 6 || slices > 8) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "slices must be 6, 7 or 8 (0 = default)");
  ctx->var_engine = engine; ctx->var_slices = slices; ctx->var_a_slices = a_slices;
  return 0;
}

int bosip_b200_get_variance_engine(bosip_b200_ctx* h, int* engine, int* slices, int* kstar_slices) {
  API_LOCK(h);
  if (engine) *engine = ctx->var_engine;
  if (slices) *slices = ctx->var_slices;
  if (kstar_slices) *kstar_slices = ctx->var_a_slices;
  return 0;
}

int bosip_b200_set_chunk(bosip_b200_ctx* h, int64_t max_points) {
  API_LOCK(h);
  if (max_points < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "chunk must be >= 0");
  ctx->user_chunk = max_points;
  return 0;
}

int bosip_b200_gp_fit(bosip_b200_ctx* h, int kernel_id, int d, int N, int p, const double* X, const double* Y,
                      const double* lambda, const double* alpha, const double* sigma, const double* mean_at_X,
                      const bosip_b200_gp_opts* opts, bosip_b200_gp** out) {
  API_LOCK(h);
  if (!out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "out is NULL");
  *out = nullptr;
  Gp* g = nullptr;
  BOSIP_TRY(gp_build(ctx, kernel_id, d, N, p, X, Y, lambda, alpha, sigma, mean_at_X, nullptr, nullptr, opts, &g));
  *out = new bosip_b200_gp{g};
  return 0;
}

int bosip_b200_gp_load_factors(bosip_b200_ctx* h, int kernel_id, int d, int N, int p, const double* X, const double* lambda,
                               const double* alpha, const double* sigma, const double* L, const double* a,
                               const bosip_b200_gp_opts* opts, bosip_b200_gp** out) {
  API_LOCK(h);
  if (!out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "out is NULL");
  *out = nullptr;
  if (!L || !a) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "L and a are required");
  Gp* g = nullptr;
  BOSIP_TRY(gp_build(ctx, kernel_id, d, N, p, X, nullptr, lambda, alpha, sigma, nullptr, L, a, opts, &g));
  *out = new bosip_b200_gp{g};
  return 0;
}

int bosip_b200_gp_get_factors(bosip_b200_gp* gh, double* L, double* Linv, double* a) {
  if (!gh || !gh->g) return BOSIP_B200_EINVAL;
  Gp* gp = gh->g;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gp->ctx);
  API_LOCK(h);
  const size_t nn = (size_t)gp->p * gp->N * gp->N * 8;
  if (L) BOSIP_CUDA(ctx, cudaMemcpy(L, gp->L, nn, cudaMemcpyDeviceToHost));
  if (Linv) BOSIP_CUDA(ctx, cudaMemcpy(Linv, gp->Linv, nn, cudaMemcpyDeviceToHost));
  if (a) BOSIP_CUDA(ctx, cudaMemcpy(a, gp->a, (size_t)gp->p * gp->N * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int bosip_b200_gp_dims(const bosip_b200_gp* gh, int* d, int* N, int* p) {
  if (!gh || !gh->g) return BOSIP_B200_EINVAL;
  if (d) *d = gh->g->d;
  if (N) *N = gh->g->N;
  if (p) *p = gh->g->p;
  return 0;
}

int bosip_b200_gp_free(bosip_b200_gp* gh) {
  if (!gh) return BOSIP_B200_EINVAL;
  if (gh->g) {
    bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gh->g->ctx);
    std::lock_guard<std::mutex> lock(h->c.mu);
    DeviceScope dev_scope(h->c.device);
    gp_destroy(gh->g);
  }
  delete gh;
  return 0;
}

int bosip_b200_gp_predict(bosip_b200_gp* gh, const double* Xq, int64_t M, int mem, const double* mean_at_Xq, double* mu,
                          double* var) {
  if (!gh || !gh->g) return BOSIP_B200_EINVAL;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gh->g->ctx);
  API_LOCK(h);
  if (!Xq && M > 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "Xq is NULL");
  if (!mu && !var) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "at least one of mu, var is required");
  Sweep sw;
  sw.gp = gh->g; sw.mem = mem; sw.xq = Xq; sw.M = M; sw.mean_at_xq = mean_at_Xq; sw.mu = mu; sw.var = var;
  sw.need_var = (var != nullptr); sw.use_like = false;
  BOSIP_TRY(make_epi(ctx, gh->g, nullptr, nullptr, &sw.ep));
  return run_sweep(ctx, sw);
}

int bosip_b200_posterior_eval_argmax(bosip_b200_gp* gh, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                     const double* Xq, int64_t M, int mem, const double* mean_at_Xq, unsigned want,
                                     double* log_approx_post, double* log_post_mean, double* log_post_var, int64_t* n_clamped,
                                     int acq, int64_t index_offset, int64_t* best_idx, double* best_val) {
  if (!gh || !gh->g) return BOSIP_B200_EINVAL;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gh->g->ctx);
  API_LOCK(h);
  if (!like) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood is NULL");
  if (!Xq && M > 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "Xq is NULL");
  const bool do_arg = acq >= 0;
  if (do_arg && acq != BOSIP_B200_ACQ_MAXVAR && acq != BOSIP_B200_ACQ_LOGMAXVAR) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown acquisition");
  if (do_arg && (!best_idx || !best_val || index_offset < 0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "argmax outputs are required");
  if ((want & 7u) == 0 && !do_arg) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "`want` selects no output");
  if (((want & BOSIP_B200_WANT_APPROX) && !log_approx_post) || ((want & BOSIP_B200_WANT_MEAN) && !log_post_mean) ||
      ((want & BOSIP_B200_WANT_VAR) && !log_post_var))
    BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "an output selected by `want` is NULL");
  Sweep sw;
  sw.gp = gh->g; sw.mem = mem; sw.xq = Xq; sw.M = M; sw.mean_at_xq = mean_at_Xq; sw.use_like = true;
  sw.la = (want & BOSIP_B200_WANT_APPROX) ? log_approx_post : nullptr;
  sw.lm = (want & BOSIP_B200_WANT_MEAN) ? log_post_mean : nullptr;
  sw.lv = (want & BOSIP_B200_WANT_VAR) ? log_post_var : nullptr;
  sw.need_var = do_arg || (want & (BOSIP_B200_WANT_MEAN | BOSIP_B200_WANT_VAR)) != 0;
  sw.do_argmax = do_arg; sw.exp_vals = (acq == BOSIP_B200_ACQ_MAXVAR) ? 1 : 0; sw.index_offset = index_offset;
  BOSIP_TRY(make_epi(ctx, gh->g, like, prior, &sw.ep));
  if (prior && prior->logprior) sw.logprior = prior->logprior;
  int rc = run_sweep(ctx, sw);
  if (n_clamped) *n_clamped = sw.n_clamped;
  if (do_arg) { *best_idx = sw.best_idx; *best_val = sw.best_idx >= 0 ? sw.best_val : -INFINITY; }
  return rc;
}

int bosip_b200_posterior_eval_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like,
                                 const bosip_b200_prior* prior, const double* Xq, int64_t M, int mem, const double* mean_at_Xq,
                                 unsigned want, double* log_approx_post, double* log_post_mean, double* log_post_var,
                                 int64_t* n_clamped, int acq, int64_t index_offset, int64_t* best_idx, double* best_val) {
  if (!gps || n_samples < 1 || !gps[0] || !gps[0]->g) return BOSIP_B200_EINVAL;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gps[0]->g->ctx);
  API_LOCK(h);
  if (!like) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "likelihood is NULL");
  if (!Xq && M > 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "Xq is NULL");
  const bool do_arg = acq >= 0;
  if (do_arg && acq != BOSIP_B200_ACQ_MAXVAR && acq != BOSIP_B200_ACQ_LOGMAXVAR) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown acquisition");
  if (do_arg && (!best_idx || !best_val || index_offset < 0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "argmax outputs are required");
  if ((want & 7u) == 0 && !do_arg) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "`want` selects no output");
  if (((want & BOSIP_B200_WANT_APPROX) && !log_approx_post) || ((want & BOSIP_B200_WANT_MEAN) && !log_post_mean) ||
      ((want & BOSIP_B200_WANT_VAR) && !log_post_var))
    BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "an output selected by `want` is NULL");
  Sweep sw;
  sw.gp = gps[0]->g; sw.mem = mem; sw.xq = Xq; sw.M = M; sw.mean_at_xq = mean_at_Xq; sw.use_like = true;
  sw.la = (want & BOSIP_B200_WANT_APPROX) ? log_approx_post : nullptr;
  sw.lm = (want & BOSIP_B200_WANT_MEAN) ? log_post_mean : nullptr;
  sw.lv = (want & BOSIP_B200_WANT_VAR) ? log_post_var : nullptr;
  sw.need_var = do_arg || (want & (BOSIP_B200_WANT_MEAN | BOSIP_B200_WANT_VAR)) != 0;
  sw.do_argmax = do_arg; sw.exp_vals = (acq == BOSIP_B200_ACQ_MAXVAR) ? 1 : 0; sw.index_offset = index_offset;
  BOSIP_TRY(set_fits(ctx, sw, gps, n_samples, like, prior));
  if (prior && prior->logprior) sw.logprior = prior->logprior;
  int rc = run_sweep(ctx, sw);
  if (n_clamped) *n_clamped = sw.n_clamped;
  if (do_arg) { *best_idx = sw.best_idx; *best_val = sw.best_idx >= 0 ? sw.best_val : -INFINITY; }
  return rc;
}

int bosip_b200_posterior_eval(bosip_b200_gp* gh, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                              const double* Xq, int64_t M, int mem, const double* mean_at_Xq, unsigned want,
                              double* log_approx_post, double* log_post_mean, double* log_post_var, int64_t* n_clamped) {
  return bosip_b200_posterior_eval_argmax(gh, like, prior, Xq, M, mem, mean_at_Xq, want, log_approx_post, log_post_mean,
                                          log_post_var, n_clamped, -1, 0, nullptr, nullptr);
}

int bosip_b200_normal_likelihood_eval(bosip_b200_ctx* h, const bosip_b200_like_normal* like, const double* mu, const double* var,
                                      int64_t M, int mem, double* marg_approx, double* l_approx, double* marg_mean,
                                      double* l_mean, double* l_sq, double* l_var, int64_t* n_clamped) {
  API_LOCK(h);
  if (!like || like->p < 1 || like->p > MAX_P) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad likelihood");
  if (M < 0 || (!mu && M > 0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "mu is required");
  if (n_clamped) *n_clamped = 0;
  if (M == 0) return 0;
  EpiParams ep; memset(&ep, 0, sizeof(ep));
  int p = like->p;  // model-output dimension of mu / var
  if (like->kind == BOSIP_B200_LIKE_NORMAL_SUM && like->sum_lengths) { p = 0; for (int i = 0; i < like->p; ++i) p += like->sum_lengths[i]; }
  if (p < 1 || p > MAX_P) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad likelihood");
  ep.p = p;
  BOSIP_TRY(make_like_params(ctx, like, p, &ep));
  const int q = ep.q;
  cudaStream_t st = ctx->s_comp;
  const bool host = mem == BOSIP_B200_HOST;
  const size_t pm = (size_t)p * M, qm = (size_t)q * M, mm = (size_t)M;
  size_t off = 0;
  auto carve = [&](size_t n) { size_t o = off; off += (n * 8 + 255) / 256 * 256; return o; };
  const size_t o_cnt = carve(4);
  size_t o_mu = 0, o_var = 0, o_ma = 0, o_la = 0, o_mm = 0, o_lm = 0, o_ls = 0, o_lv = 0;
  if (host) {
    o_mu = carve(pm); if (var) o_var = carve(pm);
    if (marg_approx) o_ma = carve(qm); if (l_approx) o_la = carve(mm); if (marg_mean) o_mm = carve(qm);
    if (l_mean) o_lm = carve(mm); if (l_sq) o_ls = carve(mm); if (l_var) o_lv = carve(mm);
  }
  BOSIP_TRY(ensure_ws(ctx, off));
  auto D = [&](size_t o) { return reinterpret_cast<double*>(ctx->ws + o); };
  unsigned long long init[2] = {0ull, ~0ull};
  BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_cnt), init, sizeof(init), cudaMemcpyHostToDevice, st));
  const double *dmu = mu, *dvar = var;
  double *dma = marg_approx, *dla = l_approx, *dmm = marg_mean, *dlm = l_mean, *dls = l_sq, *dlv = l_var;
  if (host) {
    BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_mu), mu, pm * 8, cudaMemcpyHostToDevice, st)); dmu = D(o_mu);
    if (var) { BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_var), var, pm * 8, cudaMemcpyHostToDevice, st)); dvar = D(o_var); }
    if (marg_approx) dma = D(o_ma); if (l_approx) dla = D(o_la); if (marg_mean) dmm = D(o_mm);
    if (l_mean) dlm = D(o_lm); if (l_sq) dls = D(o_ls); if (l_var) dlv = D(o_lv);
  }
  BOSIP_TRY(launch_like_eval(ctx, ep, dmu, dvar, M, dma, dla, dmm, dlm, dls, dlv, reinterpret_cast<unsigned long long*>(D(o_cnt)), st));
  if (host) {
    if (marg_approx) BOSIP_CUDA(ctx, cudaMemcpyAsync(marg_approx, dma, qm * 8, cudaMemcpyDeviceToHost, st));
    if (l_approx) BOSIP_CUDA(ctx, cudaMemcpyAsync(l_approx, dla, mm * 8, cudaMemcpyDeviceToHost, st));
    if (marg_mean) BOSIP_CUDA(ctx, cudaMemcpyAsync(marg_mean, dmm, qm * 8, cudaMemcpyDeviceToHost, st));
    if (l_mean) BOSIP_CUDA(ctx, cudaMemcpyAsync(l_mean, dlm, mm * 8, cudaMemcpyDeviceToHost, st));
    if (l_sq) BOSIP_CUDA(ctx, cudaMemcpyAsync(l_sq, dls, mm * 8, cudaMemcpyDeviceToHost, st));
    if (l_var) BOSIP_CUDA(ctx, cudaMemcpyAsync(l_var, dlv, mm * 8, cudaMemcpyDeviceToHost, st));
  }
  unsigned long long res[2];
  BOSIP_CUDA(ctx, cudaMemcpyAsync(res, D(o_cnt), sizeof(res), cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  if (n_clamped) *n_clamped = (int64_t)res[0];
  if (res[1] != ~0ull)
    BOSIP_FAIL(ctx, BOSIP_B200_EDOMAIN, "DomainError: likelihood variance below -1e-8 (src/posterior/likelihood.jl:61-65) at index " +
                                          std::to_string((long long)res[1]));
  return 0;
}

static int acq_argmax_common(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                             const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq, int64_t* best_idx,
                             double* best_val, double* best_x) {
  if (!gps || n_samples < 1 || !gps[0] || !gps[0]->g) return BOSIP_B200_EINVAL;
  bosip_b200_gp* gh = gps[0];
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gh->g->ctx);
  API_LOCK(h);
  if (!like || !src || !best_idx || !best_val) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null argument");
  if (acq != BOSIP_B200_ACQ_MAXVAR && acq != BOSIP_B200_ACQ_LOGMAXVAR) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown acquisition");
  if (index_offset < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "index_offset must be >= 0");
  Sweep sw;
  sw.M = M; sw.index_offset = index_offset; sw.use_like = true; sw.need_var = true;
  sw.do_argmax = true; sw.exp_vals = (acq == BOSIP_B200_ACQ_MAXVAR) ? 1 : 0;
  bosip_b200_cand_src shifted = *src;
  if (src->kind == BOSIP_B200_CAND_POINTS) {
    if (!src->points && M > 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "candidate points are NULL");
    sw.mem = src->mem; sw.xq = src->points;  // points are this rank's M candidates; index_offset only labels them
    if (prior && prior->logprior) sw.logprior = prior->logprior;
  } else if (src->kind == BOSIP_B200_CAND_PHILOX || src->kind == BOSIP_B200_CAND_GRID) {
    if (src->kind == BOSIP_B200_CAND_GRID && src->n_per_dim < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "n_per_dim must be >= 1");
    if (prior && prior->logprior) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "host-supplied logprior needs explicit candidate points");
    sw.mem = BOSIP_B200_DEVICE; sw.xq = nullptr; sw.src = &shifted;
  } else {
    BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown candidate source kind");
  }
  BOSIP_TRY(set_fits(ctx, sw, gps, n_samples, like, prior));
  BOSIP_TRY(run_sweep(ctx, sw));
  *best_idx = sw.best_idx;
  *best_val = sw.best_idx >= 0 ? sw.best_val : -INFINITY;
  if (best_x && sw.best_idx >= 0) {
    const int d = gh->g->d;
    const int64_t local = sw.best_idx - index_offset;
    if (src->kind == BOSIP_B200_CAND_POINTS) {
      if (src->mem == BOSIP_B200_HOST) memcpy(best_x, src->points + (size_t)local * d, (size_t)d * 8);
      else BOSIP_CUDA(ctx, cudaMemcpy(best_x, src->points + (size_t)local * d, (size_t)d * 8, cudaMemcpyDeviceToHost));
    } else {
      BOSIP_TRY(ensure_ws(ctx, 4096));
      BOSIP_TRY(launch_candidates(ctx, src, d, 1, sw.best_idx, reinterpret_cast<double*>(ctx->ws), ctx->s_comp));
      BOSIP_CUDA(ctx, cudaMemcpyAsync(best_x, ctx->ws, (size_t)d * 8, cudaMemcpyDeviceToHost, ctx->s_comp));
      BOSIP_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
    }
  }
  return 0;
}

int bosip_b200_acq_argmax(bosip_b200_gp* gh, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                          const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq, int64_t* best_idx,
                          double* best_val, double* best_x) {
  return acq_argmax_common(&gh, 1, like, prior, src, M, index_offset, acq, best_idx, best_val, best_x);
}

int bosip_b200_acq_argmax_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                             const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq, int64_t* best_idx,
                             double* best_val, double* best_x) {
  return acq_argmax_common(gps, n_samples, like, prior, src, M, index_offset, acq, best_idx, best_val, best_x);
}

int bosip_b200_marginals_int(bosip_b200_gp* gh, const bosip_b200_like_normal* like, const bosip_b200_prior* prior, const double* lhc,
                             int64_t G, int grid_size, const double* lb, const double* ub, unsigned which, double* diag, double* pairs) {
  if (!gh || !gh->g) return BOSIP_B200_EINVAL;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gh->g->ctx);
  API_LOCK(h);
  Gp* gp = gh->g;
  const int d = gp->d;
  if (!like || !lhc || !lb || !ub || G < 1 || grid_size < 2) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad arguments");
  if (which != BOSIP_B200_WANT_APPROX && which != BOSIP_B200_WANT_MEAN && which != BOSIP_B200_WANT_VAR) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "which must be one WANT_* bit");
  if (!diag && !pairs) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "no output requested");
  if (prior && prior->logprior) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "the sweep generates its points on the device: a prior descriptor is required");
  const int64_t gs = grid_size, n_diag = diag ? (int64_t)d * gs : 0, n_pairs = pairs ? (int64_t)d * (d - 1) / 2 * gs * gs : 0;
  const int64_t n_slots = n_diag + n_pairs;
  if (n_slots == 0) return 0;
  double *d_lhc = nullptr, *d_out = nullptr;
  struct Tmps { double** v[2]; ~Tmps() { for (auto q : v) if (*q) cudaFree(*q); } } tmps{{&d_lhc, &d_out}};
  BOSIP_CUDA(ctx, cudaMalloc(&d_lhc, (size_t)d * G * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&d_out, (size_t)n_slots * 8));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(d_lhc, lhc, (size_t)d * G * 8, cudaMemcpyHostToDevice, ctx->s_comp));
  bosip_b200_cand_src src;
  memset(&src, 0, sizeof(src));
  src.kind = 3; src.points = d_lhc; src.seed = (uint64_t)G; src.mem = (int32_t)n_diag; src.lb = lb; src.ub = ub; src.n_per_dim = gs;
  Sweep sw;
  sw.gp = gp; sw.mem = BOSIP_B200_DEVICE; sw.xq = nullptr; sw.src = &src; sw.M = n_slots * G; sw.index_offset = 0; sw.use_like = true;
  sw.need_var = which != BOSIP_B200_WANT_APPROX;
  sw.seg_len = G; sw.seg_out = d_out; sw.seg_which = which;
  BOSIP_TRY(make_epi(ctx, gp, like, prior, &sw.ep));
  BOSIP_TRY(run_sweep(ctx, sw));
  if (diag) BOSIP_CUDA(ctx, cudaMemcpy(diag, d_out, (size_t)n_diag * 8, cudaMemcpyDeviceToHost));
  if (pairs) BOSIP_CUDA(ctx, cudaMemcpy(pairs, d_out + n_diag, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int bosip_b200_candidates(bosip_b200_ctx* h, const bosip_b200_cand_src* src, int d, int64_t M, int64_t index_offset, int mem,
                          double* out) {
  API_LOCK(h);
  if (!src || !out || M < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad arguments");
  if (src->kind != BOSIP_B200_CAND_PHILOX && src->kind != BOSIP_B200_CAND_GRID) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "source must be PHILOX or GRID");
  if (M == 0) return 0;
  if (mem == BOSIP_B200_DEVICE) {
    BOSIP_TRY(launch_candidates(ctx, src, d, M, index_offset, out, ctx->s_comp));
  } else {
    BOSIP_TRY(ensure_ws(ctx, (size_t)M * d * 8));
    BOSIP_TRY(launch_candidates(ctx, src, d, M, index_offset, reinterpret_cast<double*>(ctx->ws), ctx->s_comp));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(out, ctx->ws, (size_t)M * d * 8, cudaMemcpyDeviceToHost, ctx->s_comp));
  }
  BOSIP_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
  return 0;
}

static int evidence_sum_common(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                               const double* xs, int64_t S, int mem, unsigned which, double* sum_out) {
  if (!gps || n_samples < 1 || !gps[0] || !gps[0]->g) return BOSIP_B200_EINVAL;
  bosip_b200_ctx* h = reinterpret_cast<bosip_b200_ctx*>(gps[0]->g->ctx);
  API_LOCK(h);
  if (!like || !prior || !xs || !sum_out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null argument");
  if (which != BOSIP_B200_WANT_APPROX && which != BOSIP_B200_WANT_MEAN) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "which must be WANT_APPROX or WANT_MEAN");
  Sweep sw;
  sw.mem = mem; sw.xq = xs; sw.M = S; sw.use_like = true;
  sw.need_var = (which == BOSIP_B200_WANT_MEAN);
  sw.do_evidence = true; sw.evid_which = which;
  BOSIP_TRY(set_fits(ctx, sw, gps, n_samples, like, prior));
  if (prior->logprior) sw.logprior = prior->logprior;
  BOSIP_TRY(run_sweep(ctx, sw));
  *sum_out = sw.evid_sum;
  return 0;
}

int bosip_b200_evidence_sum(bosip_b200_gp* gh, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                            const double* xs, int64_t S, int mem, unsigned which, double* sum_out) {
  return evidence_sum_common(&gh, 1, like, prior, xs, S, mem, which, sum_out);
}

int bosip_b200_evidence_sum_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                               const double* xs, int64_t S, int mem, unsigned which, double* sum_out) {
  return evidence_sum_common(gps, n_samples, like, prior, xs, S, mem, which, sum_out);
}

int bosip_b200_mmd_sums(bosip_b200_ctx* h, int kernel_id, int d, const double* lambda, const double* A, int64_t n,
                        const double* B, int64_t m, int mem, int rank, int world, double* sums_out) {
  API_LOCK(h);
  return mmd_sums(ctx, kernel_id, d, lambda, A, n, B, m, mem, rank, world, sums_out);
}

int bosip_b200_mmd(bosip_b200_ctx* h, int kernel_id, int d, const double* lambda, const double* A, int64_t n, const double* B,
                   int64_t m, int mem, double* mmd2_out) {
  API_LOCK(h);
  if (!mmd2_out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null output");
  double s[3];
  BOSIP_TRY(mmd_sums(ctx, kernel_id, d, lambda, A, n, B, m, mem, 0, 1, s));
  *mmd2_out = s[0] / ((double)n * (double)n) + s[1] / ((double)m * (double)m) - 2.0 * s[2] / ((double)n * (double)m);
  return 0;
}

int bosip_b200_tv_stage1(bosip_b200_ctx* h, const double* lw, const double* a, const double* t, int64_t G, int mem, double* out4) {
  API_LOCK(h);
  return tv_stage1(ctx, lw, a, t, G, mem, out4);
}
int bosip_b200_tv_stage2(bosip_b200_ctx* h, const double* lw, const double* a, const double* t, int64_t G, int mem, double eva,
                         double evt, double* out2) {
  API_LOCK(h);
  return tv_stage2(ctx, lw, a, t, G, mem, eva, evt, out2);
}
int bosip_b200_tv(bosip_b200_ctx* h, const double* lw, const double* a, const double* t, int64_t G, int mem, double* tv_out) {
  API_LOCK(h);
  if (!tv_out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null output");
  return tv_full(ctx, lw, a, t, G, mem, tv_out);
}

int bosip_b200_mvnormal_sample(bosip_b200_ctx* h, int d, const double* mu, const double* L, uint64_t seed, int64_t offset, int64_t M,
                               int mem, double* out) {
  API_LOCK(h);
  if (!mu || !L || !out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "mvnormal_sample: NULL argument");
  return mvn_sample(ctx, d, mu, L, seed, offset, M, mem, out);
}

int bosip_b200_mixture_pdf_sum(bosip_b200_ctx* h, int d, int nq, const double* mus, const double* Linvs, const double* lognorm,
                               const double* X, int64_t M, int mem, int accumulate, double* delta) {
  API_LOCK(h);
  if (!mus || !Linvs || !lognorm || !X || !delta) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "mixture_pdf_sum: NULL argument");
  return mix_pdf_sum(ctx, d, nq, mus, Linvs, lognorm, X, M, mem, accumulate, delta);
}

int bosip_b200_amis_log_weights(bosip_b200_ctx* h, const double* log_p, const double* delta, double t, int64_t M, int mem, double* log_w) {
  API_LOCK(h);
  if (!log_p || !delta || !log_w) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "amis_log_weights: NULL argument");
  return amis_log_weights(ctx, log_p, delta, t, M, mem, log_w);
}

int bosip_b200_weighted_moments(bosip_b200_ctx* h, int d, const double* X, const double* log_w, int64_t M, int mem, double* mean,
                                double* cov, double* sums, double* ws) {
  API_LOCK(h);
  if (!X || !log_w || !mean || !cov) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "weighted_moments: NULL argument");
  return weighted_moments(ctx, d, X, log_w, M, mem, mean, cov, sums, ws);
}

int bosip_b200_resample(bosip_b200_ctx* h, int d, const double* X, const double* ws, int64_t M, uint64_t seed, int64_t count, int mem,
                        double* out, int64_t* idx) {
  API_LOCK(h);
  if (!X || !ws || !out) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "resample: NULL argument");
  return resample(ctx, d, X, ws, M, seed, count, mem, out, idx);
}

int bosip_b200_weighted_quantile(bosip_b200_ctx* h, const double* vals, const double* ws, int64_t n, int mem, double p, double* out) {
  API_LOCK(h);
  return cs_weighted_quantile(ctx, vals, ws, n, mem, p, out);
}
int bosip_b200_cutoff_area(bosip_b200_ctx* h, const double* vals, const double* ws, int64_t n, int mem, double c, double* out) {
  API_LOCK(h);
  return cs_cutoff_area(ctx, vals, ws, n, mem, c, out);
}
int bosip_b200_set_iou(bosip_b200_ctx* h, const uint8_t* in_A, const uint8_t* in_B, const double* logprior, int64_t n, int mem, double* out) {
  API_LOCK(h);
  return cs_set_iou(ctx, in_A, in_B, logprior, n, mem, out);
}
int bosip_b200_confidence_iou(bosip_b200_ctx* h, const double* log_fa, const double* log_fb, const double* logprior, int64_t n, int mem,
                              double q, double* iou, double* c_a, double* c_b) {
  API_LOCK(h);
  return cs_confidence_iou(ctx, log_fa, log_fb, logprior, n, mem, q, iou, c_a, c_b);
}

int bosip_b200_int8_peak(bosip_b200_ctx* h, double* burst_tops, double* sustained_tops) {
  API_LOCK(h);
  if (!burst_tops) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "burst_tops must not be NULL");
  return i8_peak(ctx, burst_tops, sustained_tops, nullptr);
}

int bosip_b200_int8_peak_random(bosip_b200_ctx* h, double* sustained_tops) {
  API_LOCK(h);
  if (!sustained_tops) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "sustained_tops must not be NULL");
  double burst = 0.0;
  return i8_peak(ctx, &burst, nullptr, sustained_tops);
}

int bosip_b200_fp64_peak(bosip_b200_ctx* h, double* dmma, double* dfma) {
  API_LOCK(h);
  return fp64_peak(ctx, dmma, dfma);
}

int bosip_b200_math_selftest(bosip_b200_ctx* h, const double* x, int64_t n, int variant, double* ef, double* qf, double* er, double* qr) {
  API_LOCK(h);
  if (!x || !ef || !qf || !er || !qr) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null argument");
  return math_selftest(ctx, x, n, variant, ef, qf, er, qr);
}

int bosip_b200_timing_reset(bosip_b200_ctx* h, int enable) {
  API_LOCK(h);
  ctx->timing = Timing();
  ctx->timing.enabled = enable != 0;
  return 0;
}
int bosip_b200_timing_get(bosip_b200_ctx* h, double* ms3, int64_t* launches3) {
  API_LOCK(h);
  for (int k = 0; k < 3; ++k) { if (ms3) ms3[k] = ctx->timing.ms[k]; if (launches3) launches3[k] = ctx->timing.launches[k]; }
  return 0;
}

}  // extern "C"
