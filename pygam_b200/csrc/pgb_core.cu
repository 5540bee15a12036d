// pgb_core.cu -- model descriptor, column statistics, the model-matrix and interval passes (the statistics /
// predict pass is in pgb_pass.cu, the Gram pass in pgb_gram.cu and pgb_cells.cu).
//
// Reference code replaced (pyGAM 0.12.0): GAM._pirls steps pygam.py:764-783,
// GAM._linear_predictor 385-421, _estimate_model_statistics 1031-1056 (data part),
// _initial_estimate 696-710 (data part), TermList.build_columns terms.py:1901-1923,
// gen_edge_knots utils.py:592-621.  See DESIGN.md for the kernel design.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "pgb_internal.h"
#include "pgb_kernels.cuh"

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


// ----------------------------------------------------------------------------------------
// model descriptor
// ----------------------------------------------------------------------------------------
static int term_nnz(const pgb_term& t) {
  if (t.kind == PGB_TERM_INTERCEPT || t.kind == PGB_TERM_LINEAR) return 1;
  int nnz = 1;
  for (int q = 0; q < t.n_marg; ++q) nnz *= (t.order[q] < 0) ? 1 : t.order[q] + 1;  // order -1: a linear marginal of te()
  return nnz;
}

int pgb_plan_gram(pgb_model* M);      // pgb_gram.cu
int pgb_plan_upload(pgb_model* M);    // pgb_gram.cu
int pgb_plan_coverage_errors(const pgb_model* M);
void pgb_gram_release(pgb_model* M);  // pgb_gram.cu
int pgb_plan_cells(pgb_model* M);     // pgb_cells.cu
int pgb_cells_upload(pgb_model* M);
int pgb_cells_coverage_errors(const pgb_model* M);
void pgb_cells_release(pgb_model* M);

int pgb_model_build_host(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist,
                            int32_t link, double levels, double expectile, int32_t device, const pgb_debug_opts* dbg,
                            pgb_model** out) {
  if (!terms || n_terms < 1 || !out) { pgb_set_error("pgb_model_create: bad arguments"); return PGB_EINVAL; }
  if (dist < 0 || dist > PGB_DIST_INVGAUSS || link < 0 || link > PGB_LINK_INVSQUARED) {
    pgb_set_error("unknown distribution %d or link %d", dist, link);
    return PGB_EINVAL;
  }
  pgb_model* M = new pgb_model();
  if (dbg) {
    M->dbg.force_general = dbg->force_general_path;
    M->dbg.cells_tile_rows = dbg->cells_tile_rows;
    M->dbg.cells_band_tiles = dbg->cells_band_tiles;
    M->dbg.cells_threads = dbg->cells_threads;
    M->dbg.cells_kernel = dbg->cells_kernel;
  }
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
      if (t.kind == PGB_TERM_TENSOR && g.order == PGB_MARG_LINEAR) {
        // a LinearTerm marginal of te(): one column, the feature value itself (terms.py:693-708 inside 1454-1477)
        ok = ok && g.nspl == 1 && !g.periodic;
        g.nint = 1;
        g.lo = 0.0;
        g.span = g.inv_span = g.knot_step = 1.0;
        continue;
      }
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
    d.fast = (t.kind == PGB_TERM_SPLINE && t.order[0] == 3 && !t.periodic[0] && t.by < 0) ? 1 : 0;
    d.fc0 = -1;
    d.slot0 = slot;
    slot += d.nnz;
    coef += t.n_coefs;
    M->terms.push_back(t);
    M->dterms.push_back(d);
  }
  M->m = coef;
  M->k = slot;
  if (M->m > 65000) { pgb_set_error("too many coefficients (%d)", M->m); delete M; return PGB_EUNSUPPORTED; }
  // The general Gram path keeps a whole term-pair block of G in one SM's shared memory: a 200-spline s() or a 25 x 25
  // te() does not fit.  Such a model is fine as long as the cell-owned pass takes it (plain cubic s(), te() of two
  // cubics, f(), l()): the general plan is then simply absent.
  int rc = pgb_plan_gram(M);
  char general_err[512] = "";
  if (rc == PGB_EUNSUPPORTED && !M->dbg.force_general) {
    snprintf(general_err, sizeof(general_err), "%s", pgb_last_error());
    M->general_ok = false;
    M->roles.clear();
    M->elem_map.clear();
    M->gram_ctas = 0;
    M->partial_doubles = 0;
    rc = PGB_OK;
  }
  if (rc == PGB_OK) rc = pgb_plan_cells(M);
  if (rc == PGB_OK && !M->general_ok && !M->cells.enabled) {
    pgb_set_error("%s, and the model has terms the cell-owned Gram pass does not take (by=, cyclic or non-cubic bases, "
                  "te() of more than two marginals)", general_err);
    rc = PGB_EUNSUPPORTED;
  }
  if (rc != PGB_OK) { delete M; return rc; }
  *out = M;
  return PGB_OK;
}

static int model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist, int32_t link,
                        double levels, double expectile, int32_t device, const pgb_debug_opts* dbg, pgb_model** out) {
  pgb_model* M = nullptr;
  int rc = pgb_model_build_host(terms, n_terms, n_features, dist, link, levels, expectile, device, dbg, &M);
  if (rc != PGB_OK) return rc;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&M->d_terms, sizeof(DTerm) * n_terms);
  if (e == cudaSuccess)
    e = cudaMemcpy(M->d_terms, M->dterms.data(), sizeof(DTerm) * n_terms, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = pgb_cuda_fail(e, "pgb_model_create");
  if (rc == PGB_OK) rc = pgb_plan_upload(M);
  if (rc == PGB_OK) rc = pgb_cells_upload(M);
  if (rc != PGB_OK) {
    cudaFree(M->d_terms);
    pgb_gram_release(M);
    pgb_cells_release(M);
    delete M;
    return rc;
  }
  *out = M;
  return PGB_OK;
}

extern "C" int pgb_model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features,
                                int32_t dist, int32_t link, double levels, double expectile,
                                int32_t device, pgb_model** out) {
  return model_create(terms, n_terms, n_features, dist, link, levels, expectile, device, nullptr, out);
}

// test-only entry point: the same model with the plan of its Gram pass constrained (see pgb_debug_opts)
extern "C" int pgb_debug_model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist,
                                      int32_t link, double levels, double expectile, int32_t device,
                                      const pgb_debug_opts* opts, pgb_model** out) {
  return model_create(terms, n_terms, n_features, dist, link, levels, expectile, device, opts, out);
}

// Host-only planning of the Gram pass for a model descriptor (no CUDA call): fills `info` and
// returns in *coverage_errors how many entries of the upper triangle of [G | r] are not produced
// by exactly one accumulator element (0 for a valid plan).  Used by the CPU test-suite.
extern "C" int pgb_model_plan_host(const pgb_term* terms, int32_t n_terms, int32_t n_features,
                                   pgb_model_info* info, int32_t* coverage_errors) {
  pgb_model* M = nullptr;
  int rc = pgb_model_build_host(terms, n_terms, n_features, PGB_DIST_NORMAL, PGB_LINK_IDENTITY, 1.0, -1.0, 0, nullptr, &M);
  if (rc != PGB_OK) return rc;
  if (info) pgb_model_get_info(M, info);
  if (coverage_errors) *coverage_errors = (M->general_ok ? pgb_plan_coverage_errors(M) : 0) + pgb_cells_coverage_errors(M);
  delete M;
  return PGB_OK;
}

extern "C" void pgb_model_destroy(pgb_model* M) {
  if (!M) return;
  pgb_gram_release(M);
  pgb_cells_release(M);
  cudaFree(M->d_terms);
  delete M;
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
  info->gram_ctas = M->gram_ctas;
  info->gram_ws_bytes = pgb_gram_ws_bytes(M, 0);
  info->pass_ws_bytes = pgb_stats_ws_bytes(M);
  info->cell_roles = M->cells.enabled ? (int32_t)(M->cells.roles.size() + M->cells.groles.size()) : 0;
  info->cell_tile_rows = M->cells.enabled ? M->cells.tile_rows : 0;
  info->cell_threads = M->cells.enabled ? M->cells.threads : 0;
  info->cell_ring_groups = M->cells.enabled ? M->cells.ring_groups : 0;
  info->cell_generic_roles = M->cells.enabled ? (int32_t)M->cells.groles.size() : 0;
  info->reserved0 = 0;
  return PGB_OK;
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
// interval pass: per row the linear predictor of one term (or the whole model) and its variance
// x_i^T cov x_i with the basis row evaluated on the fly -- GAM._get_quantiles (pygam.py:1399-1404)
// over _modelmat / _linear_predictor restricted to a term (pygam.py:385-421, 469-497)
// ----------------------------------------------------------------------------------------
static const int kIntervalMaxNnz = 320;
struct EmitCollect {
  int32_t* col;
  double* val;
  int n;
  __device__ __forceinline__ void operator()(int c, double v) {
    if (n < kIntervalMaxNnz) { col[n] = c; val[n] = v; }
    ++n;
  }
};
__global__ void __launch_bounds__(128)
interval_pass_kernel(const DTerm* __restrict__ terms, int n_terms, int m, int term, const double* __restrict__ X,
                     int64_t n, int64_t ld, const double* __restrict__ beta, const double* __restrict__ cov,
                     double* __restrict__ lp_out, double* __restrict__ var_out) {
  int32_t col[kIntervalMaxNnz];
  double val[kIntervalMaxNnz];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    EmitCollect e{col, val, 0};
    if (term < 0) {
      for (int t = 0; t < n_terms; ++t) eval_term(terms[t], X, ld, i, e);
    } else {
      eval_term(terms[term], X, ld, i, e);
    }
    const int k = min(e.n, kIntervalMaxNnz);
    double lp = 0.0, var = 0.0;
    for (int a = 0; a < k; ++a) lp = fma(val[a], __ldg(beta + col[a]), lp);
    if (var_out) {
      for (int a = 0; a < k; ++a) {
        const double* crow = cov + (size_t)col[a] * m;
        double s = 0.5 * val[a] * __ldg(crow + col[a]);
        for (int b = a + 1; b < k; ++b) s = fma(val[b], __ldg(crow + col[b]), s);
        var = fma(2.0 * val[a], s, var);  // cov is symmetric: diagonal once, every off-diagonal pair twice
      }
    }
    if (lp_out) lp_out[i] = lp;
    if (var_out) var_out[i] = var;
  }
}
extern "C" int pgb_interval_pass(const pgb_model* M, const double* X, int64_t n, int64_t ld, const double* beta,
                                 const double* cov, int32_t term, double* lp_out, double* var_out, void* stream) {
  if (!M || !X || !beta || n < 1 || ld < n || term < -1 || term >= M->n_terms || (var_out && !cov)) {
    pgb_set_error("pgb_interval_pass: bad arguments");
    return PGB_EINVAL;
  }
  int k = 0;
  for (int t = 0; t < M->n_terms; ++t)
    if (term < 0 || t == term) k += M->dterms[t].nnz;
  if (k > kIntervalMaxNnz) {
    pgb_set_error("pgb_interval_pass: %d non-zeros per row (limit %d)", k, kIntervalMaxNnz);
    return PGB_EUNSUPPORTED;
  }
  const int grid = (int)std::min<int64_t>((n + 127) / 128, kSMs * 8);
  interval_pass_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(M->d_terms, M->n_terms, M->m, term, X, n, ld, beta, cov,
                                                                lp_out, var_out);
  PGB_LAUNCHED(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}
