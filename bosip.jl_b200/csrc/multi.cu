// Multi-GPU behind the C ABI: ONE host process drives N B200s through N per-device contexts.
//
// A Julia host that reaches this library through `BOSS.maximize_acquisition` (/root/reference/src/main.jl:98-101,
// docs/src/hyperparams.md:69) or through the posterior closures (src/posterior/log_posterior.jl:15-61) has no torchrun and
// no torch.distributed: it makes one blocking ccall.  The `_multi` entry points shard that call exactly as parallel.py does
// for the one-process-per-GPU launch -- contiguous ranges of the query/candidate points with global indices preserved, the
// MMD tile grid dealt round-robin, the TV grid cut into contiguous slices -- run the shards concurrently on one host thread
// per device, and combine the few doubles that come back in FIXED RANK ORDER on the calling thread:
//   argmax            per-device (value, global index, x) -> Julia `argmax` order, lowest global index on ties
//   evidence / MMD    per-device partial sums added in rank order
//   TV                per-device (max, sum) log-sum-exp pairs merged in rank order between the two stages
// The fit is replicated: every device conditions the same inputs (N^3/3 is tiny next to one sweep; identical GPUs give
// bit-identical factors), so nothing heavier than those doubles ever crosses between devices and no NCCL communicator is
// needed.  Per-point outputs and the argmax are bit-identical to a one-GPU call (every point's arithmetic is independent
// of the chunking); sums differ from the one-GPU sum only in the association of the per-shard partials.
#include <math.h>
#include <string.h>

#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bosip_b200.h"

struct bosip_b200_multi {
  std::vector<bosip_b200_ctx*> ctx;
  std::vector<int> dev;
  std::string err;
};
struct bosip_b200_gp_multi {
  bosip_b200_multi* mh;
  std::vector<bosip_b200_gp*> gp;
};

namespace {

thread_local std::string g_multi_init_error;

// contiguous shard of `total` items: the first total % world ranks get one extra (parallel.shard_range)
void shard_range(int64_t total, int rank, int world, int64_t* start, int64_t* count) {
  const int64_t base = total / world, rem = total % world;
  *start = rank * base + (rank < rem ? rank : rem);
  *count = base + (rank < rem ? 1 : 0);
}

// run fn(rank) on one host thread per device; returns the first non-zero code in rank order and records its message
int run_all(bosip_b200_multi* mh, const std::function<int(int)>& fn) {
  const int n = (int)mh->ctx.size();
  std::vector<int> rc(n, 0);
  if (n == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(n);
    for (int r = 0; r < n; ++r) th.emplace_back([&, r] { rc[r] = fn(r); });
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < n; ++r)
    if (rc[r] != 0) {
      mh->err = "device " + std::to_string(mh->dev[r]) + ": " + bosip_b200_last_error(mh->ctx[r]);
      return rc[r];
    }
  return 0;
}

struct Lse { double m, s; };
Lse lse_merge(Lse a, Lse b) {   // same -Inf-safe merge as metrics.cu / parallel.lse_merge
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  Lse r;
  if (a.m >= b.m) { r.m = a.m; r.s = (a.m == INFINITY) ? a.s : a.s + b.s * exp(b.m - a.m); }
  else { r.m = b.m; r.s = (b.m == INFINITY) ? b.s : b.s + a.s * exp(a.m - b.m); }
  return r;
}
double lme(Lse x, double G) {   // logmeanexp from a (max, sum) pair; src/utils/utils.jl:7-15
  if (isinf(x.m) || x.s <= 0.0) return isinf(x.m) ? x.m : -INFINITY;
  return x.m + log(x.s) - log(G);
}

// Julia argmax order over per-rank winners (idx < 0: that rank had no candidate)
int combine_argmax(const std::vector<double>& v, const std::vector<int64_t>& idx) {
  int best = -1;
  for (int r = 0; r < (int)v.size(); ++r) {
    if (idx[r] < 0) continue;
    if (best < 0) { best = r; continue; }
    const bool an = v[best] != v[best], bn = v[r] != v[r];
    if (an || bn) { if (bn && (!an || idx[r] < idx[best])) best = r; }
    else if (v[r] > v[best] || (v[r] == v[best] && idx[r] < idx[best])) best = r;
  }
  return best;
}

}  // namespace

extern "C" {

int bosip_b200_init_multi(int ndev, const int* devs, bosip_b200_multi** out) {
  if (!out || ndev < 1) return BOSIP_B200_EINVAL;
  *out = nullptr;
  bosip_b200_multi* mh = new bosip_b200_multi();
  for (int r = 0; r < ndev; ++r) {
    const int dev = devs ? devs[r] : r;
    bosip_b200_ctx* c = nullptr;
    const int rc = bosip_b200_init(dev, &c);
    if (rc != 0) {
      g_multi_init_error = std::string("device ") + std::to_string(dev) + ": " + bosip_b200_last_error(nullptr);
      for (auto* q : mh->ctx) bosip_b200_shutdown(q);
      delete mh;
      return rc;
    }
    mh->ctx.push_back(c); mh->dev.push_back(dev);
  }
  *out = mh;
  return 0;
}

int bosip_b200_shutdown_multi(bosip_b200_multi* mh) {
  if (!mh) return BOSIP_B200_EINVAL;
  for (auto* c : mh->ctx) bosip_b200_shutdown(c);
  delete mh;
  return 0;
}

int bosip_b200_multi_size(const bosip_b200_multi* mh) { return mh ? (int)mh->ctx.size() : 0; }

bosip_b200_ctx* bosip_b200_multi_context(bosip_b200_multi* mh, int rank) {
  return (mh && rank >= 0 && rank < (int)mh->ctx.size()) ? mh->ctx[rank] : nullptr;
}

const char* bosip_b200_multi_last_error(const bosip_b200_multi* mh) { return mh ? mh->err.c_str() : g_multi_init_error.c_str(); }

int bosip_b200_gp_fit_multi(bosip_b200_multi* mh, int kernel_id, int d, int N, int p, const double* X, const double* Y,
                            const double* lambda, const double* alpha, const double* sigma, const double* mean_at_X,
                            const bosip_b200_gp_opts* opts, bosip_b200_gp_multi** out) {
  if (!mh || !out) return BOSIP_B200_EINVAL;
  *out = nullptr;
  bosip_b200_gp_multi* g = new bosip_b200_gp_multi();
  g->mh = mh; g->gp.assign(mh->ctx.size(), nullptr);
  const int rc = run_all(mh, [&](int r) {
    return bosip_b200_gp_fit(mh->ctx[r], kernel_id, d, N, p, X, Y, lambda, alpha, sigma, mean_at_X, opts, &g->gp[r]);
  });
  if (rc != 0) {
    for (auto* q : g->gp) if (q) bosip_b200_gp_free(q);
    delete g;
    return rc;
  }
  *out = g;
  return 0;
}

bosip_b200_gp* bosip_b200_gp_multi_replica(bosip_b200_gp_multi* g, int rank) {
  return (g && rank >= 0 && rank < (int)g->gp.size()) ? g->gp[rank] : nullptr;
}

int bosip_b200_gp_free_multi(bosip_b200_gp_multi* g) {
  if (!g) return BOSIP_B200_EINVAL;
  for (auto* q : g->gp) if (q) bosip_b200_gp_free(q);
  delete g;
  return 0;
}

int bosip_b200_posterior_eval_multi(bosip_b200_gp_multi* g, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                    const double* Xq, int64_t M, const double* mean_at_Xq, unsigned want, double* log_approx_post,
                                    double* log_post_mean, double* log_post_var, int64_t* n_clamped, int acq, int64_t index_offset,
                                    int64_t* best_idx, double* best_val) {
  if (!g || !g->mh) return BOSIP_B200_EINVAL;
  bosip_b200_multi* mh = g->mh;
  const int n = (int)g->gp.size();
  if (M < 0 || (!Xq && M > 0)) { mh->err = "Xq is NULL"; return BOSIP_B200_EINVAL; }
  if (acq >= 0 && (!best_idx || !best_val)) { mh->err = "argmax outputs are required"; return BOSIP_B200_EINVAL; }
  std::vector<int64_t> ncl(n, 0), bi(n, -1);
  std::vector<double> bv(n, 0.0);
  int dim = 0, pdim = 0;   // strides of the d x M queries and the p x M prior-mean values
  if (bosip_b200_gp_dims(g->gp[0], &dim, nullptr, &pdim) != 0) { mh->err = "bad GP handle"; return BOSIP_B200_EINVAL; }
  const int rc = run_all(mh, [&](int r) {
    int64_t s0, cnt;
    shard_range(M, r, n, &s0, &cnt);
    if (cnt == 0) return 0;
    bosip_b200_prior pr;
    const bosip_b200_prior* prp = prior;
    if (prior && prior->logprior) { pr = *prior; pr.logprior = prior->logprior + s0; prp = &pr; }
    return bosip_b200_posterior_eval_argmax(g->gp[r], like, prp, Xq + (size_t)s0 * dim, cnt, BOSIP_B200_HOST,
                                            mean_at_Xq ? mean_at_Xq + (size_t)s0 * pdim : nullptr, want,
                                            log_approx_post ? log_approx_post + s0 : nullptr, log_post_mean ? log_post_mean + s0 : nullptr,
                                            log_post_var ? log_post_var + s0 : nullptr, &ncl[r], acq, index_offset + s0,
                                            acq >= 0 ? &bi[r] : nullptr, acq >= 0 ? &bv[r] : nullptr);
  });
  if (rc != 0) return rc;
  if (n_clamped) { *n_clamped = 0; for (int r = 0; r < n; ++r) *n_clamped += ncl[r]; }
  if (acq >= 0) {
    const int w = combine_argmax(bv, bi);
    *best_idx = w < 0 ? -1 : bi[w];
    *best_val = w < 0 ? -INFINITY : bv[w];
  }
  return 0;
}

int bosip_b200_acq_argmax_multi(bosip_b200_gp_multi* g, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq, int64_t* best_idx,
                                double* best_val, double* best_x) {
  if (!g || !g->mh) return BOSIP_B200_EINVAL;
  bosip_b200_multi* mh = g->mh;
  const int n = (int)g->gp.size();
  if (!src || !best_idx || !best_val || M < 0) { mh->err = "null argument"; return BOSIP_B200_EINVAL; }
  if (src->kind == BOSIP_B200_CAND_POINTS && src->mem != BOSIP_B200_HOST) {
    mh->err = "acq_argmax_multi: explicit candidate points must be HOST memory (a device pointer belongs to one GPU)";
    return BOSIP_B200_EINVAL;
  }
  int dim = 0;
  if (bosip_b200_gp_dims(g->gp[0], &dim, nullptr, nullptr) != 0) { mh->err = "bad GP handle"; return BOSIP_B200_EINVAL; }
  std::vector<int64_t> bi(n, -1);
  std::vector<double> bv(n, 0.0), bx((size_t)n * dim, 0.0);
  const int rc = run_all(mh, [&](int r) {
    int64_t s0, cnt;
    shard_range(M, r, n, &s0, &cnt);
    if (cnt == 0) return 0;
    bosip_b200_cand_src sr = *src;
    bosip_b200_prior pr;
    const bosip_b200_prior* prp = prior;
    if (src->kind == BOSIP_B200_CAND_POINTS) {
      sr.points = src->points + (size_t)s0 * dim;
      if (prior && prior->logprior) { pr = *prior; pr.logprior = prior->logprior + s0; prp = &pr; }
    }
    return bosip_b200_acq_argmax(g->gp[r], like, prp, &sr, cnt, index_offset + s0, acq, &bi[r], &bv[r], &bx[(size_t)r * dim]);
  });
  if (rc != 0) return rc;
  const int w = combine_argmax(bv, bi);
  *best_idx = w < 0 ? -1 : bi[w];
  *best_val = w < 0 ? -INFINITY : bv[w];
  if (best_x && w >= 0) memcpy(best_x, &bx[(size_t)w * dim], (size_t)dim * 8);
  return 0;
}

int bosip_b200_evidence_sum_multi(bosip_b200_gp_multi* g, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                  const double* xs, int64_t S, unsigned which, double* sum_out) {
  if (!g || !g->mh) return BOSIP_B200_EINVAL;
  bosip_b200_multi* mh = g->mh;
  const int n = (int)g->gp.size();
  if (!xs || !sum_out || !prior || S < 0) { mh->err = "null argument"; return BOSIP_B200_EINVAL; }
  int dim = 0;
  if (bosip_b200_gp_dims(g->gp[0], &dim, nullptr, nullptr) != 0) { mh->err = "bad GP handle"; return BOSIP_B200_EINVAL; }
  std::vector<double> part(n, 0.0);
  const int rc = run_all(mh, [&](int r) {
    int64_t s0, cnt;
    shard_range(S, r, n, &s0, &cnt);
    if (cnt == 0) return 0;
    bosip_b200_prior pr = *prior;
    if (prior->logprior) pr.logprior = prior->logprior + s0;
    return bosip_b200_evidence_sum(g->gp[r], like, &pr, xs + (size_t)s0 * dim, cnt, BOSIP_B200_HOST, which, &part[r]);
  });
  if (rc != 0) return rc;
  double s = 0.0;
  for (int r = 0; r < n; ++r) s += part[r];   // rank order
  *sum_out = s;
  return 0;
}

int bosip_b200_mmd_multi(bosip_b200_multi* mh, int kernel_id, int d, const double* lambda, const double* A, int64_t na, const double* B,
                         int64_t nb, double* mmd2_out) {
  if (!mh || !mmd2_out) return BOSIP_B200_EINVAL;
  const int n = (int)mh->ctx.size();
  std::vector<double> sums((size_t)n * 3, 0.0);
  const int rc = run_all(mh, [&](int r) {
    return bosip_b200_mmd_sums(mh->ctx[r], kernel_id, d, lambda, A, na, B, nb, BOSIP_B200_HOST, r, n, &sums[(size_t)r * 3]);
  });
  if (rc != 0) return rc;
  double s[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < 3; ++c) s[c] += sums[(size_t)r * 3 + c];
  *mmd2_out = s[0] / ((double)na * (double)na) + s[1] / ((double)nb * (double)nb) - 2.0 * s[2] / ((double)na * (double)nb);
  return 0;
}

int bosip_b200_tv_multi(bosip_b200_multi* mh, const double* lw, const double* a, const double* t, int64_t G, double* tv_out) {
  if (!mh || !tv_out) return BOSIP_B200_EINVAL;
  const int n = (int)mh->ctx.size();
  if (!lw || !a || !t || G < 1) { mh->err = "bad TV arguments"; return BOSIP_B200_EINVAL; }
  if (n == 1) return bosip_b200_tv(mh->ctx[0], lw, a, t, G, BOSIP_B200_HOST, tv_out);
  std::vector<double> s1((size_t)n * 4, 0.0), s2((size_t)n * 2, 0.0);
  std::vector<char> has(n, 0);
  int rc = run_all(mh, [&](int r) {
    int64_t s0, cnt;
    shard_range(G, r, n, &s0, &cnt);
    if (cnt == 0) return 0;
    has[r] = 1;
    return bosip_b200_tv_stage1(mh->ctx[r], lw + s0, a + s0, t + s0, cnt, BOSIP_B200_HOST, &s1[(size_t)r * 4]);
  });
  if (rc != 0) return rc;
  Lse la{-INFINITY, 0.0}, lt{-INFINITY, 0.0};
  for (int r = 0; r < n; ++r)
    if (has[r]) { la = lse_merge(la, Lse{s1[(size_t)r * 4], s1[(size_t)r * 4 + 1]}); lt = lse_merge(lt, Lse{s1[(size_t)r * 4 + 2], s1[(size_t)r * 4 + 3]}); }
  const double eva = lme(la, (double)G), evt = lme(lt, (double)G);
  rc = run_all(mh, [&](int r) {
    int64_t s0, cnt;
    shard_range(G, r, n, &s0, &cnt);
    if (cnt == 0) return 0;
    return bosip_b200_tv_stage2(mh->ctx[r], lw + s0, a + s0, t + s0, cnt, BOSIP_B200_HOST, eva, evt, &s2[(size_t)r * 2]);
  });
  if (rc != 0) return rc;
  Lse ld{-INFINITY, 0.0};
  for (int r = 0; r < n; ++r)
    if (has[r]) ld = lse_merge(ld, Lse{s2[(size_t)r * 2], s2[(size_t)r * 2 + 1]});
  *tv_out = 0.5 * exp(lme(ld, (double)G));
  return 0;
}

}  // extern "C"
