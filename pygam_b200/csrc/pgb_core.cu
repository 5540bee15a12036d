// pgb_core.cu -- model descriptor, column statistics, the O(k)-per-row passes (model matrix,
// statistics / predict) and the fused Gram pass of the PIRLS iteration.
//
// Reference code replaced (pyGAM 0.12.0): GAM._pirls steps pygam.py:764-783,
// GAM._linear_predictor 385-421, _estimate_model_statistics 1031-1056 (data part),
// _initial_estimate 696-710 (data part), TermList.build_columns terms.py:1901-1923,
// gen_edge_knots utils.py:592-621.  See DESIGN.md for the kernel design.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "pgb_internal.h"

// ----------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void pgb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int pgb_cuda_fail(cudaError_t e, const char* what) {
  pgb_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return PGB_ECUDA;
}
extern "C" const char* pgb_last_error(void) { return g_err; }
extern "C" int pgb_version(void) { return 100; }

#include <atomic>
static std::atomic<long long> g_launches{0};
#define PGB_LAUNCHED(n) g_launches.fetch_add((n), std::memory_order_relaxed)
extern "C" long long pgb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void pgb_count_launch(int n) { PGB_LAUNCHED(n); }

static const int kSMs = 148;                 // B200
static const size_t kSmemMax = 227 * 1024;   // opt-in shared memory per CTA on sm_100
static const int kPassThreads = 256;

// ----------------------------------------------------------------------------------------
// model descriptor
// ----------------------------------------------------------------------------------------
static int term_nnz(const pgb_term& t) {
  if (t.kind == PGB_TERM_INTERCEPT || t.kind == PGB_TERM_LINEAR) return 1;
  int nnz = 1;
  for (int q = 0; q < t.n_marg; ++q) nnz *= t.order[q] + 1;
  return nnz;
}

static void stage_free(pgb_model* M);

static int plan_gram(pgb_model* M) {
  const int m = M->m, k = M->k;
  // shared memory: [strip][rec_val T*kq][rec_w T][rec_col T*kq u16][terms][beta m]
  const size_t fixed = (size_t)M->n_terms * sizeof(DTerm) + (size_t)m * 8 + PGB_GS_COUNT * 32 * 8 + 512;
  const size_t rec_row = (size_t)(k + 1) * 10 + 8;
  int T = 256;
  while (T > 32 && (size_t)T * rec_row > 96 * 1024) T >>= 1;
  if ((size_t)T * rec_row + fixed + 16 * 1024 > kSmemMax) {
    pgb_set_error("model too wide for the Gram pass: k=%d non-zeros per row", k);
    return PGB_EUNSUPPORTED;
  }
  M->tile_rows = T;
  const size_t strip_budget = (kSmemMax - fixed - (size_t)T * rec_row) / 8;  // doubles
  // whole upper triangle in one role?  then try to leave room for 2 CTAs per SM
  M->roles.clear();
  int rlo = 0;
  int64_t off = 0;
  while (rlo < m) {
    const int ldp = (m - rlo + 1) | 1;
    int rows_fit = (int)std::min<int64_t>((int64_t)(strip_budget / ldp), m - rlo);
    if (rows_fit < 1) {
      pgb_set_error("Gram strip does not fit in shared memory (m=%d)", m);
      return PGB_EUNSUPPORTED;
    }
    int rhi = rlo + rows_fit;
    if (rhi < m) {
      // prefer a term boundary if one exists inside (rlo, rhi]
      int best = -1;
      for (const DTerm& t : M->dterms) {
        const int end = t.coef_offset + t.n_coefs;
        if (end > rlo && end <= rhi) best = std::max(best, end);
      }
      if (best > rlo) rhi = best;
    }
    DRole r{};
    r.rlo = rlo;
    r.rhi = rhi;
    r.ldp = ldp;
    r.first_term = 0;
    for (int t = 0; t < M->n_terms; ++t) {
      if (M->dterms[t].coef_offset + M->dterms[t].n_coefs > rlo) { r.first_term = t; break; }
    }
    r.slot0 = M->dterms[r.first_term].slot0;
    r.kq = k - r.slot0 + 1;
    int last_slot = r.slot0;
    for (int t = r.first_term; t < M->n_terms; ++t) {
      if (M->dterms[t].coef_offset < rhi) last_slot = M->dterms[t].slot0 + M->dterms[t].nnz;
    }
    r.row_slots = last_slot - r.slot0;
    r.strip_off = off;
    r.strip_size = (int64_t)(rhi - rlo) * ldp;
    off += r.strip_size;
    M->roles.push_back(r);
    rlo = rhi;
  }
  M->partial_stride = off;
  // balance roles: equalise strip sizes a little by re-splitting evenly when > 1 role
  size_t smem = 0;
  for (const DRole& r : M->roles) {
    const size_t s = (size_t)r.strip_size * 8 + (size_t)T * ((size_t)r.kq * 10 + 8) + fixed;
    smem = std::max(smem, s);
  }
  M->gram_smem = (smem + 127) & ~(size_t)127;
  M->occupancy = (int)std::max<size_t>(1, std::min<size_t>(4, kSmemMax / (M->gram_smem + 1024)));
  const int n_roles = (int)M->roles.size();
  M->gram_ctas_per_role = std::max(1, (kSMs * M->occupancy) / n_roles);
  M->pass_ctas = kSMs * 8;
  return PGB_OK;
}

extern "C" int pgb_model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features,
                                int32_t dist, int32_t link, double levels, double expectile,
                                int32_t device, pgb_model** out) {
  if (!terms || n_terms < 1 || !out) { pgb_set_error("pgb_model_create: bad arguments"); return PGB_EINVAL; }
  if (dist < 0 || dist > PGB_DIST_INVGAUSS || link < 0 || link > PGB_LINK_INVSQUARED) {
    pgb_set_error("unknown distribution %d or link %d", dist, link);
    return PGB_EINVAL;
  }
  pgb_model* M = new pgb_model();
  M->device = device;
  M->n_terms = n_terms;
  M->n_features = n_features;
  M->fam.dist = dist;
  M->fam.link = link;
  M->fam.levels = levels;
  M->fam.expectile = expectile;
  int slot = 0, coef = 0;
  for (int i = 0; i < n_terms; ++i) {
    const pgb_term& t = terms[i];
    DTerm d{};
    d.kind = t.kind;
    d.n_marg = t.n_marg;
    d.by = t.by;
    d.drop_first = t.drop_first;
    d.coef_offset = t.coef_offset;
    d.n_coefs = t.n_coefs;
    bool ok = t.kind >= PGB_TERM_INTERCEPT && t.kind <= PGB_TERM_TENSOR && t.coef_offset == coef &&
              t.n_coefs >= 1 && t.by < n_features;
    const int nm = (t.kind == PGB_TERM_INTERCEPT) ? 0 : t.n_marg;
    if (t.kind == PGB_TERM_TENSOR) ok = ok && nm >= 2 && nm <= PGB_MAX_MARG;
    else if (t.kind != PGB_TERM_INTERCEPT) ok = ok && nm == 1;
    int ncoef = 1;
    for (int q = 0; q < nm && ok; ++q) {
      DMarg& g = d.marg[q];
      g.feature = t.feature[q];
      g.nspl = t.n_splines[q];
      g.order = t.order[q];
      g.periodic = t.periodic[q];
      ok = ok && g.feature >= 0 && g.feature < n_features;
      if (t.kind == PGB_TERM_LINEAR) continue;
      ok = ok && g.order >= 0 && g.order <= PGB_MAX_ORDER && g.nspl > g.order;
      g.nint = g.nspl + (g.periodic ? g.order : 0) - g.order;  // utils.py:678, 686
      g.lo = std::min(t.edge_lo[q], t.edge_hi[q]);             // np.sort(edge_knots), utils.py:681
      g.span = std::max(t.edge_lo[q], t.edge_hi[q]) - g.lo;
      if (g.span == 0.0) g.span = 1.0;                          // utils.py:684-685
      g.inv_span = 1.0 / g.span;
      g.knot_step = 1.0 / (double)g.nint;
      ncoef *= g.nspl;
    }
    if (t.kind == PGB_TERM_FACTOR && t.drop_first) ncoef -= 1;
    ok = ok && ncoef == t.n_coefs;
    if (!ok) {
      pgb_set_error("pgb_model_create: term %d is inconsistent", i);
      delete M;
      return PGB_EINVAL;
    }
    d.nnz = term_nnz(t);
    d.slot0 = slot;
    slot += d.nnz;
    coef += t.n_coefs;
    M->terms.push_back(t);
    M->dterms.push_back(d);
  }
  M->m = coef;
  M->k = slot;
  if (M->m > 65000) { pgb_set_error("too many coefficients (%d)", M->m); delete M; return PGB_EUNSUPPORTED; }
  int rc = plan_gram(M);
  if (rc != PGB_OK) { delete M; return rc; }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&M->d_terms, sizeof(DTerm) * n_terms);
  if (e == cudaSuccess) e = cudaMalloc(&M->d_roles, sizeof(DRole) * M->roles.size());
  if (e == cudaSuccess)
    e = cudaMemcpy(M->d_terms, M->dterms.data(), sizeof(DTerm) * n_terms, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(M->d_roles, M->roles.data(), sizeof(DRole) * M->roles.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    rc = pgb_cuda_fail(e, "pgb_model_create");
    cudaFree(M->d_terms);
    cudaFree(M->d_roles);
    delete M;
    return rc;
  }
  *out = M;
  return PGB_OK;
}

extern "C" void pgb_model_destroy(pgb_model* M) {
  if (!M) return;
  stage_free(M);
  if (M->timing.ev[0]) { cudaEventDestroy(M->timing.ev[0]); cudaEventDestroy(M->timing.ev[1]); }
  cudaFree(M->d_terms);
  cudaFree(M->d_roles);
  delete M;
}

static int64_t gram_fixed_ws_doubles(const pgb_model* M) {
  const int64_t ctas = (int64_t)M->gram_ctas_per_role * (int64_t)M->roles.size();
  return (int64_t)M->gram_ctas_per_role * M->partial_stride + (ctas + M->pass_ctas) * PGB_GS_COUNT;
}

extern "C" int64_t pgb_gram_ws_bytes(const pgb_model* M, int64_t n) {
  if (!M) return 0;
  int64_t d = gram_fixed_ws_doubles(M);
  if (M->roles.size() > 1) d += 2 * n;  // (omega, z) per row
  return d * 8 + 256;
}
extern "C" int64_t pgb_stats_ws_bytes(const pgb_model* M) {
  return M ? (int64_t)M->pass_ctas * PGB_SS_COUNT * 8 + 256 : 0;
}

extern "C" int pgb_model_get_info(const pgb_model* M, pgb_model_info* info) {
  if (!M || !info) { pgb_set_error("pgb_model_get_info: null"); return PGB_EINVAL; }
  info->m = M->m;
  info->k = M->k;
  info->n_features = M->n_features;
  info->n_roles = (int32_t)M->roles.size();
  info->tile_rows = M->tile_rows;
  info->gram_ctas = M->gram_ctas_per_role * (int32_t)M->roles.size();
  info->gram_ws_bytes = gram_fixed_ws_doubles(M) * 8 + 256;
  info->pass_ws_bytes = pgb_stats_ws_bytes(M);
  return PGB_OK;
}

// ----------------------------------------------------------------------------------------
// deterministic block reduction of S per-thread sums -> dst[0..S)
// ----------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void block_reduce_store(double (&acc)[S], double* red /* S*32 smem */,
                                                   double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    double v = acc[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[s * 32 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < S) {
    double v = 0.0;
    for (int w = 0; w < nw; ++w) v += red[threadIdx.x * 32 + w];
    dst[threadIdx.x] = v;
  }
}

// sum `count` partial vectors of length S in a fixed order: out[s] (+)= sum_c part[c*S+s]
__global__ void scalar_reduce_kernel(const double* __restrict__ part, int count, int S,
                                     double* __restrict__ out, int accumulate) {
  const int s = threadIdx.x;
  if (s >= S) return;
  double v = accumulate ? out[s] : 0.0;
  for (int c = 0; c < count; ++c) v += part[(int64_t)c * S + s];
  out[s] = v;
}

// ----------------------------------------------------------------------------------------
// column statistics: gen_edge_knots (utils.py:592-621) + the finite check of check_X
// ----------------------------------------------------------------------------------------
__global__ void col_stats_kernel(const double* __restrict__ X, int64_t n, int64_t ld,
                                 double* __restrict__ part /* [F][gridDim.x][3] */) {
  const int f = blockIdx.y;
  const double* x = X + (int64_t)f * ld;
  double lo = INFINITY, hi = -INFINITY, bad = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ldg(x + i);
    if (isfinite(v)) { lo = fmin(lo, v); hi = fmax(hi, v); }
    else bad += 1.0;
  }
  __shared__ double s_lo[32], s_hi[32], s_bad[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
    bad += __shfl_down_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; s_bad[warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); bad += s_bad[w];
    }
    double* p = part + ((int64_t)f * gridDim.x + blockIdx.x) * 3;
    p[0] = lo; p[1] = hi; p[2] = bad;
  }
}
__global__ void col_stats_final_kernel(const double* __restrict__ part, int nblk, int F,
                                       double* __restrict__ minmax, long long* __restrict__ nonfinite) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double lo = INFINITY, hi = -INFINITY, bad = 0.0;
  for (int b = 0; b < nblk; ++b) {
    const double* p = part + ((int64_t)f * nblk + b) * 3;
    lo = fmin(lo, p[0]); hi = fmax(hi, p[1]); bad += p[2];
  }
  minmax[2 * f] = lo;
  minmax[2 * f + 1] = hi;
  nonfinite[f] = (long long)bad;
}

extern "C" int pgb_col_stats(const double* X, int64_t n, int64_t ld, int32_t F, double* minmax,
                             long long* nonfinite, void* ws, int64_t ws_bytes, void* stream) {
  const int nblk = 296;
  if (!X || n < 1 || F < 1 || ws_bytes < (int64_t)F * nblk * 24) {
    pgb_set_error("pgb_col_stats: bad arguments (workspace needs %lld bytes)", (long long)F * nblk * 24);
    return PGB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  col_stats_kernel<<<dim3(nblk, F), 256, 0, st>>>(X, n, ld, (double*)ws);
  col_stats_final_kernel<<<(F + 63) / 64, 64, 0, st>>>((const double*)ws, nblk, F, minmax, nonfinite);
  PGB_LAUNCHED(2);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}
extern "C" int64_t pgb_col_stats_ws_bytes(int32_t F) { return (int64_t)F * 296 * 24; }

__global__ void col_levels_kernel(const double* __restrict__ x, int64_t n, double lo, int range,
                                  int* __restrict__ present) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ldg(x + i) - lo;
    if (v >= 0.0 && v < (double)range) {
      const int b = (int)v;
      if ((double)b == v) { if (present[b] == 0) present[b] = 1; }
      else present[range] = 1;  // a non-integer level: the caller falls back to an exact count
    }
  }
}
extern "C" int pgb_col_levels(const double* x, int64_t n, double lo, int32_t range, int32_t* present,
                              void* stream) {
  if (!x || !present || range < 1) { pgb_set_error("pgb_col_levels: bad arguments"); return PGB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  PGB_CUDA(cudaMemsetAsync(present, 0, sizeof(int) * (range + 1), st));
  col_levels_kernel<<<kSMs * 4, 256, 0, st>>>(x, n, lo, range, present);
  PGB_LAUNCHED(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// ----------------------------------------------------------------------------------------
// dense model matrix (inspection / tests): TermList.build_columns, terms.py:1901-1923
// ----------------------------------------------------------------------------------------
struct EmitDense {
  double* row;
  __device__ __forceinline__ void operator()(int c, double v) { row[c] += v; }
};
__global__ void model_matrix_kernel(const DTerm* __restrict__ terms, int n_terms, int m,
                                    const double* __restrict__ X, int64_t n, int64_t ld,
                                    double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double* row = out + i * m;
    for (int c = 0; c < m; ++c) row[c] = 0.0;
    EmitDense e{row};
    for (int t = 0; t < n_terms; ++t) eval_term(terms[t], X, ld, i, e);
  }
}
extern "C" int pgb_model_matrix(const pgb_model* M, const double* X, int64_t n, int64_t ld,
                                double* out, void* stream) {
  if (!M || !X || !out || n < 1 || ld < n) { pgb_set_error("pgb_model_matrix: bad arguments"); return PGB_EINVAL; }
  const int grid = (int)std::min<int64_t>((n + 127) / 128, kSMs * 8);
  model_matrix_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(M->d_terms, M->n_terms, M->m, X, n, ld, out);
  PGB_LAUNCHED(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// ----------------------------------------------------------------------------------------
// statistics / predict pass: lp, mu and the sums of _estimate_model_statistics on ALL rows
// ----------------------------------------------------------------------------------------
struct EmitLp {
  const double* beta;
  double lp;
  __device__ __forceinline__ void operator()(int c, double v) { lp = fma(v, beta[c], lp); }
};

__device__ __forceinline__ void load_terms_smem(DTerm* s_terms, const DTerm* g_terms, int n_terms) {
  const int words = n_terms * (int)(sizeof(DTerm) / 4);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(g_terms);
  uint32_t* dst = reinterpret_cast<uint32_t*>(s_terms);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(kPassThreads)
stats_pass_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam,
                  const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta, double scale,
                  double ybar, int poisson_round, double* __restrict__ part,
                  double* __restrict__ lp_out, double* __restrict__ mu_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_beta + ((m + 1) & ~1));
  double* s_red = reinterpret_cast<double*>(s_terms + n_terms);
  for (int i = threadIdx.x; i < m; i += blockDim.x) s_beta[i] = beta[i];
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();

  double acc[PGB_SS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_SS_COUNT; ++s) acc[s] = 0.0;
  const double null_mu = ybar;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    EmitLp e{s_beta, 0.0};
    for (int t = 0; t < n_terms; ++t) eval_term(s_terms[t], X, ld, i, e);
    const double lp = e.lp;
    const double mu = link_inv(fam, lp);
    if (lp_out) lp_out[i] = lp;
    if (mu_out) mu_out[i] = mu;
    if (y) {
      const double yi = __ldg(y + i);
      const double wi = w ? (double)__ldg(w + i) : 1.0;
      const double d = dist_dev(fam, yi, mu);
      const double r = yi - mu;
      acc[PGB_SS_N] += 1.0;
      acc[PGB_SS_WDEV] += wi * d;
      acc[PGB_SS_PEARSON] += wi * (r * r) / dist_var(fam, mu);  // distributions.py:78
      acc[PGB_SS_DEV] += d;
      acc[PGB_SS_NCORRECT] += (yi == ((mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      acc[PGB_SS_SUMY] += yi;
      acc[PGB_SS_SUMW] += wi;
      if (!isfinite(lp) || !isfinite(mu)) acc[PGB_SS_NONFINITE] += 1.0;
      if (scale > 0.0) {
        const double yl = poisson_round ? rint(yi * wi) : yi;  // pygam.py:2797-2798
        acc[PGB_SS_LOGLIK] += dist_logpdf(fam, yl, mu, wi, scale);
        acc[PGB_SS_NULL_WDEV] += wi * dist_dev(fam, yi, null_mu);
        acc[PGB_SS_NULL_LOGLIK] += dist_logpdf(fam, yl, null_mu, wi, scale);
      }
    } else if (!isfinite(lp) || !isfinite(mu)) {
      acc[PGB_SS_NONFINITE] += 1.0;
    }
  }
  block_reduce_store<PGB_SS_COUNT>(acc, s_red, part + (int64_t)blockIdx.x * PGB_SS_COUNT);
}

extern "C" int pgb_stats_pass(const pgb_model* M, const double* X, int64_t n, int64_t ld,
                              const double* y, const float* w, const double* beta, double scale,
                              double ybar, int32_t poisson_round, double* scalars, double* lp_out,
                              double* mu_out, void* ws, int64_t ws_bytes, void* stream) {
  if (!M || !X || !beta || n < 1 || ld < n) { pgb_set_error("pgb_stats_pass: bad arguments"); return PGB_EINVAL; }
  if (ws_bytes < pgb_stats_ws_bytes(M) || !ws) { pgb_set_error("pgb_stats_pass: workspace too small"); return PGB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((n + kPassThreads - 1) / kPassThreads, M->pass_ctas);
  const size_t smem = (size_t)((M->m + 1) & ~1) * 8 + (size_t)M->n_terms * sizeof(DTerm) + PGB_SS_COUNT * 32 * 8;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PGB_CUDA(cudaFuncSetAttribute(stats_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    attr_set = true;
  }
  stats_pass_kernel<<<grid, kPassThreads, smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, X, n, ld, y, w,
                                                      beta, scale, ybar, poisson_round, (double*)ws,
                                                      lp_out, mu_out);
  PGB_LAUNCHED(1);
  PGB_CUDA(cudaGetLastError());
  if (scalars) {
    scalar_reduce_kernel<<<1, 32, 0, st>>>((const double*)ws, grid, PGB_SS_COUNT, scalars, 0);
    PGB_LAUNCHED(1);
    PGB_CUDA(cudaGetLastError());
  }
  return PGB_OK;
}

// ----------------------------------------------------------------------------------------
// Gram pass
// ----------------------------------------------------------------------------------------
// per-row IRLS quantities when the Gram matrix is split over several roles: every role's
// CTAs read (omega, z) instead of re-evaluating the full linear predictor
__global__ void __launch_bounds__(kPassThreads)
irls_rows_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam, int mode,
                 const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                 const float* __restrict__ w, const double* __restrict__ beta,
                 double* __restrict__ wz, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_beta + ((m + 1) & ~1));
  double* s_red = reinterpret_cast<double*>(s_terms + n_terms);
  for (int i = threadIdx.x; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();
  double acc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) acc[s] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yi = __ldg(y + i);
    double om, z;
    if (mode == PGB_MODE_IRLS) {
      EmitLp e{s_beta, 0.0};
      for (int t = 0; t < n_terms; ++t) eval_term(s_terms[t], X, ld, i, e);
      const IrlsRow r = irls_row(fam, e.lp, yi, weight_inv(w, i));
      om = r.omega;
      z = r.z;
      if (r.keep) {
        acc[PGB_GS_NUSED] += 1.0;
        acc[PGB_GS_DEV] += dist_dev(fam, yi, r.mu);
        acc[PGB_GS_NCORRECT] += (yi == ((r.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      om = 1.0;
      z = init_response(fam, yi);
      acc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(z)) acc[PGB_GS_NONFINITE] += 1.0;
    }
    acc[PGB_GS_NROWS] += 1.0;
    wz[2 * i] = om;
    wz[2 * i + 1] = z;
  }
  block_reduce_store<PGB_GS_COUNT>(acc, s_red, part + (int64_t)blockIdx.x * PGB_GS_COUNT);
}

struct EmitRec {
  uint16_t* col;
  double* val;
  const double* beta;
  double lp;
  int q;
  __device__ __forceinline__ void operator()(int c, double v) {
    col[q] = (uint16_t)c;
    val[q] = v;
    ++q;
    if (beta) lp = fma(v, beta[c], lp);
  }
};

// grid (chunks, roles).  Each CTA walks its rows in tiles of T: phase A evaluates the basis of
// T rows into shared-memory row records (column, value) and the IRLS weight; phase B adds the
// weighted outer products of the records into the role's strip of G (shared memory, fp64),
// each warp owning the strip rows with (row & 7) == warp so that no update needs an atomic.
// The strip is written to this CTA's partial at the end; gram_reduce_kernel sums partials in a
// fixed order, which makes the whole pass bit-reproducible.
__global__ void __launch_bounds__(256)
gram_accum_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m,
                  const DRole* __restrict__ g_roles, Family fam, int mode, int inline_irls,
                  const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta,
                  const double* __restrict__ wz, int64_t rows_per_chunk, int T,
                  double* __restrict__ partials, int64_t partial_stride, double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DRole role = g_roles[blockIdx.y];
  const int kq = role.kq;
  double* strip = reinterpret_cast<double*>(smem_raw);
  double* rec_val = strip + role.strip_size;
  double* rec_w = rec_val + (size_t)T * kq;
  double* s_beta = rec_w + T;
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + PGB_GS_COUNT * 32);
  uint16_t* rec_col = reinterpret_cast<uint16_t*>(s_terms + n_terms);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < role.strip_size; i += blockDim.x) strip[i] = 0.0;
  if (inline_irls) for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();

  double acc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) acc[s] = 0.0;

  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t row_end = min(n, row_begin + rows_per_chunk);
  const int rlo = role.rlo, rhi = role.rhi, ldp = role.ldp, row_slots = role.row_slots;
  const int npass = (kq + 31) >> 5;

  for (int64_t tile = row_begin; tile < row_end; tile += T) {
    // ---------------- phase A: row records ----------------
    for (int rI = tid; rI < T; rI += blockDim.x) {
      const int64_t i = tile + rI;
      double om = 0.0, z = 0.0;
      uint16_t* rc = rec_col + (size_t)rI * kq;
      double* rv = rec_val + (size_t)rI * kq;
      if (i < row_end) {
        EmitRec e{rc, rv, inline_irls ? s_beta : nullptr, 0.0, 0};
        for (int t = role.first_term; t < n_terms; ++t) eval_term(s_terms[t], X, ld, i, e);
        if (inline_irls) {
          const double yi = __ldg(y + i);
          if (mode == PGB_MODE_IRLS) {
            const IrlsRow r = irls_row(fam, e.lp, yi, weight_inv(w, i));
            om = r.omega;
            z = r.z;
            if (r.keep) {
              acc[PGB_GS_NUSED] += 1.0;
              acc[PGB_GS_DEV] += dist_dev(fam, yi, r.mu);
              acc[PGB_GS_NCORRECT] += (yi == ((r.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
            }
          } else {
            om = 1.0;
            z = init_response(fam, yi);
            acc[PGB_GS_NUSED] += 1.0;
            if (!isfinite(z)) acc[PGB_GS_NONFINITE] += 1.0;
          }
          acc[PGB_GS_NROWS] += 1.0;
        } else {
          om = __ldg(wz + 2 * i);
          z = __ldg(wz + 2 * i + 1);
        }
      }
      rc[kq - 1] = (uint16_t)m;  // virtual column: right-hand side r = sum omega * z * b
      rv[kq - 1] = z;
      rec_w[rI] = om;
    }
    __syncthreads();
    // ---------------- phase B: strip += omega * b b^T ----------------
    const int tile_n = (int)min((int64_t)T, row_end - tile);
    for (int rI = 0; rI < tile_n; ++rI) {
      const double om = rec_w[rI];
      if (om == 0.0) continue;
      const uint16_t* rc = rec_col + (size_t)rI * kq;
      const double* rv = rec_val + (size_t)rI * kq;
      for (int qb = 0; qb < row_slots; qb += 32) {
        const int q = qb + lane;
        const int cq = (q < row_slots) ? (int)rc[q] : 0xffff;
        const bool own = (cq >= rlo) && (cq < rhi) && ((cq & 7) == warp);
        unsigned mask = __ballot_sync(0xffffffffu, own);
        while (mask) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          const int cqb = __shfl_sync(0xffffffffu, cq, b);
          const double wq = om * rv[qb + b];
          double* grow = strip + (size_t)(cqb - rlo) * ldp - rlo;
          for (int p = 0; p < npass; ++p) {
            const int qc = (p << 5) + lane;
            if (qc < kq) {
              const int cv = rc[qc];
              if (cv >= cqb) grow[cv] = fma(wq, rv[qc], grow[cv]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---------------- write the partial ----------------
  double* dst = partials + (int64_t)blockIdx.x * partial_stride + role.strip_off;
  for (int64_t i = tid; i < role.strip_size; i += blockDim.x) dst[i] = strip[i];
  if (inline_irls) {
    block_reduce_store<PGB_GS_COUNT>(acc, s_red, scal_part + (int64_t)blockIdx.x * PGB_GS_COUNT);
  }
}

// out = [G (m*m, symmetric) | r (m)]: sum the chunk partials in order and mirror the triangle
__global__ void gram_reduce_kernel(const double* __restrict__ partials, int n_chunks,
                                   int64_t partial_stride, const DRole* __restrict__ roles,
                                   int n_roles, int m, double* __restrict__ out, int accumulate) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= partial_stride) return;
  int r = 0;
  while (r + 1 < n_roles && e >= roles[r + 1].strip_off) ++r;
  const DRole role = roles[r];
  const int64_t loc = e - role.strip_off;
  const int row = role.rlo + (int)(loc / role.ldp);
  const int col = role.rlo + (int)(loc % role.ldp);
  if (col > m || col < row) return;
  double v = 0.0;
  for (int c = 0; c < n_chunks; ++c) v += partials[(int64_t)c * partial_stride + e];
  if (col == m) {
    double* dst = out + (int64_t)m * m + row;
    *dst = accumulate ? *dst + v : v;
  } else {
    double* a = out + (int64_t)row * m + col;
    const double nv = accumulate ? *a + v : v;
    *a = nv;
    out[(int64_t)col * m + row] = nv;
  }
}

static int launch_gram(const pgb_model* M, int32_t mode, const double* X, int64_t n, int64_t ld,
                       const double* y, const float* w, const double* beta, double* out, void* ws,
                       int64_t ws_bytes, int accumulate, cudaStream_t st) {
  if (!M || !X || !y || !out || !ws || n < 1 || ld < n || (mode == PGB_MODE_IRLS && !beta) ||
      (mode != PGB_MODE_IRLS && mode != PGB_MODE_INIT)) {
    pgb_set_error("pgb_gram_pass: bad arguments");
    return PGB_EINVAL;
  }
  if (ws_bytes < pgb_gram_ws_bytes(M, n)) {
    pgb_set_error("pgb_gram_pass: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)pgb_gram_ws_bytes(M, n));
    return PGB_EINVAL;
  }
  const int n_roles = (int)M->roles.size();
  const int T = M->tile_rows;
  int n_chunks = M->gram_ctas_per_role;
  int64_t rows_per_chunk = (n + n_chunks - 1) / n_chunks;
  rows_per_chunk = ((rows_per_chunk + T - 1) / T) * T;
  n_chunks = (int)((n + rows_per_chunk - 1) / rows_per_chunk);
  double* partials = (double*)ws;
  double* scal_part = partials + (int64_t)M->gram_ctas_per_role * M->partial_stride;
  double* wz = scal_part + ((int64_t)M->gram_ctas_per_role * n_roles + M->pass_ctas) * PGB_GS_COUNT;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PGB_CUDA(cudaFuncSetAttribute(gram_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    PGB_CUDA(cudaFuncSetAttribute(irls_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    attr_set = true;
  }
  const int inline_irls = (n_roles == 1);
  int scal_count = n_chunks;
  if (!inline_irls) {
    const int grid = (int)std::min<int64_t>((n + kPassThreads - 1) / kPassThreads, M->pass_ctas);
    const size_t smem = (size_t)((M->m + 1) & ~1) * 8 + (size_t)M->n_terms * sizeof(DTerm) + PGB_GS_COUNT * 32 * 8;
    irls_rows_kernel<<<grid, kPassThreads, smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, mode, X, n, ld, y,
                                                       w, beta, wz, scal_part);
    PGB_LAUNCHED(1);
    PGB_CUDA(cudaGetLastError());
    scal_count = grid;
  }
  pgb_model::Timing& tm = const_cast<pgb_model*>(M)->timing;
  if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.ev[0], st));
  gram_accum_kernel<<<dim3(n_chunks, n_roles), M->gram_threads, M->gram_smem, st>>>(
      M->d_terms, M->n_terms, M->m, M->d_roles, M->fam, mode, inline_irls, X, n, ld, y, w, beta, wz,
      rows_per_chunk, T, partials, M->partial_stride, scal_part);
  if (tm.enabled) { PGB_CUDA(cudaEventRecord(tm.ev[1], st)); tm.pending = true; }
  PGB_LAUNCHED(3);
  PGB_CUDA(cudaGetLastError());
  const int rb = 256;
  gram_reduce_kernel<<<(unsigned)((M->partial_stride + rb - 1) / rb), rb, 0, st>>>(
      partials, n_chunks, M->partial_stride, M->d_roles, n_roles, M->m, out, accumulate);
  scalar_reduce_kernel<<<1, 32, 0, st>>>(scal_part, scal_count, PGB_GS_COUNT,
                                         out + (int64_t)M->m * M->m + M->m, accumulate);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// optional timing of the dominant kernel (gram_accum_kernel) with CUDA events recorded on the
// launching stream; used by bench.py for the roofline figure
extern "C" int pgb_gram_timing(pgb_model* M, int32_t enable) {
  if (!M) { pgb_set_error("pgb_gram_timing: null model"); return PGB_EINVAL; }
  pgb_model::Timing& tm = M->timing;
  if (enable && !tm.ev[0]) {
    PGB_CUDA(cudaEventCreate(&tm.ev[0]));
    PGB_CUDA(cudaEventCreate(&tm.ev[1]));
  }
  tm.enabled = enable != 0;
  tm.pending = false;
  return PGB_OK;
}
extern "C" int pgb_gram_last_ms(pgb_model* M, float* accum_ms) {
  if (!M || !accum_ms || !M->timing.pending) { pgb_set_error("pgb_gram_last_ms: no timed pass"); return PGB_EINVAL; }
  PGB_CUDA(cudaEventSynchronize(M->timing.ev[1]));
  PGB_CUDA(cudaEventElapsedTime(accum_ms, M->timing.ev[0], M->timing.ev[1]));
  return PGB_OK;
}

extern "C" int pgb_gram_pass(const pgb_model* M, int32_t mode, const double* X, int64_t n, int64_t ld,
                             const double* y, const float* w, const double* beta, double* out, void* ws,
                             int64_t ws_bytes, void* stream) {
  return launch_gram(M, mode, X, n, ld, y, w, beta, out, ws, ws_bytes, 0, (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------------------
// host-buffer entry point: stream row chunks through two device staging buffers so that the
// host->device copies overlap the Gram kernels (the end-to-end path of bench.py)
// ----------------------------------------------------------------------------------------
static void stage_free(pgb_model* M) {
  pgb_model::HostStage& S = M->stage;
  for (int q = 0; q < 2; ++q) {
    if (S.st[q]) cudaStreamSynchronize(S.st[q]);
    cudaFree(S.dX[q]); cudaFree(S.dy[q]); cudaFree(S.dw[q]); cudaFree(S.ws[q]);
    if (S.done[q]) cudaEventDestroy(S.done[q]);
    if (S.st[q]) cudaStreamDestroy(S.st[q]);
  }
  cudaFree(S.d_out);
  cudaFree(S.d_beta);
  S = pgb_model::HostStage();
}

static int stage_ensure(pgb_model* M, int64_t chunk_rows, bool need_w) {
  pgb_model::HostStage& S = M->stage;
  if (S.cap_rows >= chunk_rows && (S.has_w || !need_w)) return PGB_OK;
  stage_free(M);
  const int F = M->n_features, m = M->m;
  cudaError_t e = cudaSuccess;
  auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  S.ws_bytes = pgb_gram_ws_bytes(M, chunk_rows);
  for (int b = 0; b < 2; ++b) {
    ck(cudaStreamCreateWithFlags(&S.st[b], cudaStreamNonBlocking));
    ck(cudaEventCreateWithFlags(&S.done[b], cudaEventDisableTiming));
    ck(cudaMalloc(&S.dX[b], sizeof(double) * chunk_rows * F));
    ck(cudaMalloc(&S.dy[b], sizeof(double) * chunk_rows));
    if (need_w) ck(cudaMalloc(&S.dw[b], sizeof(float) * chunk_rows));
    ck(cudaMalloc(&S.ws[b], S.ws_bytes));
  }
  ck(cudaMalloc(&S.d_out, sizeof(double) * ((int64_t)m * m + m + PGB_GS_COUNT)));
  ck(cudaMalloc(&S.d_beta, sizeof(double) * m));
  if (e != cudaSuccess) {
    stage_free(M);
    return pgb_cuda_fail(e, "pgb_gram_pass_host: staging allocation");
  }
  S.cap_rows = chunk_rows;
  S.has_w = need_w;
  return PGB_OK;
}

extern "C" int pgb_gram_pass_host(pgb_model* M, int32_t mode, const double* X_host, int64_t n,
                                  int64_t ld, const double* y_host, const float* w_host,
                                  const double* beta_host, double* out_host, int64_t chunk_rows) {
  if (!M || !X_host || !y_host || !out_host || n < 1 || ld < n || (mode == PGB_MODE_IRLS && !beta_host)) {
    pgb_set_error("pgb_gram_pass_host: bad arguments");
    return PGB_EINVAL;
  }
  if (chunk_rows <= 0) chunk_rows = 1 << 20;
  chunk_rows = std::min(chunk_rows, n);
  const int F = M->n_features, m = M->m;
  const int64_t out_doubles = (int64_t)m * m + m + PGB_GS_COUNT;
  PGB_CUDA(cudaSetDevice(M->device));
  int rc = stage_ensure(M, chunk_rows, w_host != nullptr);
  if (rc != PGB_OK) return rc;
  pgb_model::HostStage& S = M->stage;
  chunk_rows = S.cap_rows >= n ? std::min(S.cap_rows, n) : chunk_rows;
  cudaError_t e = cudaSuccess;
  auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  ck(cudaMemsetAsync(S.d_out, 0, sizeof(double) * out_doubles, S.st[0]));
  if (beta_host) ck(cudaMemcpyAsync(S.d_beta, beta_host, sizeof(double) * m, cudaMemcpyHostToDevice, S.st[0]));
  ck(cudaEventRecord(S.done[0], S.st[0]));
  ck(cudaEventRecord(S.done[1], S.st[1]));
  ck(cudaStreamWaitEvent(S.st[1], S.done[0], 0));
  int b = 0;
  for (int64_t r0 = 0; r0 < n && e == cudaSuccess && rc == PGB_OK; r0 += chunk_rows, b ^= 1) {
    const int64_t nr = std::min(chunk_rows, n - r0);
    // copies of this buffer may overlap the other buffer's kernels ...
    for (int f = 0; f < F; ++f)
      ck(cudaMemcpyAsync(S.dX[b] + (int64_t)f * S.cap_rows, X_host + (int64_t)f * ld + r0, sizeof(double) * nr,
                         cudaMemcpyHostToDevice, S.st[b]));
    ck(cudaMemcpyAsync(S.dy[b], y_host + r0, sizeof(double) * nr, cudaMemcpyHostToDevice, S.st[b]));
    if (w_host) ck(cudaMemcpyAsync(S.dw[b], w_host + r0, sizeof(float) * nr, cudaMemcpyHostToDevice, S.st[b]));
    // ... but the accumulation into d_out is ordered chunk by chunk
    ck(cudaStreamWaitEvent(S.st[b], S.done[b ^ 1], 0));
    if (e != cudaSuccess) break;
    rc = launch_gram(M, mode, S.dX[b], nr, S.cap_rows, S.dy[b], w_host ? S.dw[b] : nullptr, S.d_beta, S.d_out,
                     S.ws[b], S.ws_bytes, 1, S.st[b]);
    ck(cudaEventRecord(S.done[b], S.st[b]));
  }
  if (e == cudaSuccess && rc == PGB_OK) {
    ck(cudaStreamSynchronize(S.st[0]));
    ck(cudaStreamSynchronize(S.st[1]));
    ck(cudaMemcpy(out_host, S.d_out, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost));
  }
  if (e != cudaSuccess) return pgb_cuda_fail(e, "pgb_gram_pass_host");
  return rc;
}
