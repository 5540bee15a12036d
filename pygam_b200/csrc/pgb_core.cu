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
  for (int q = 0; q < t.n_marg; ++q) nnz *= t.order[q] + 1;
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
  int rc = pgb_plan_gram(M);
  if (rc == PGB_OK) rc = pgb_plan_cells(M);
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
  if (coverage_errors) *coverage_errors = pgb_plan_coverage_errors(M) + pgb_cells_coverage_errors(M);
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

// ----------------------------------------------------------------------------------------
// statistics / predict pass: lp, mu and the sums of _estimate_model_statistics on ALL rows
// (_linear_predictor pygam.py:385-421, predict_mu 423-450, the n-sized sums of 1031-1056)
// ----------------------------------------------------------------------------------------
// An O(k)-per-row pass: 8 F + 8 bytes per row against ~20 fp64-pipe slots per loaded double
// (B200: 62.7 fp64 FMA/clk/SM vs 6.5 TB/s), so both the instruction count per row and the bytes in
// flight matter.
//  * plain cubic s() terms (the `fast` flag) do not evaluate their four basis functions: every CTA
//    turns beta into one cubic polynomial per knot interval (s_tab, four coefficients per interval)
//    and a row costs an interval search and a Horner step -- ten fp64 instructions per term instead
//    of ~26.  What bounds the pass is the gather of those 32 bytes per (row, term) from shared
//    memory: the rows of a warp sit in unrelated intervals, and a plain table costs 17 wavefronts per
//    warp and term where 8 move the bytes (ncu, round 1).  The table is therefore stored as two
//    16-byte halves per interval, (c0, c1) and (c2, c3), in `rep` = 8 copies: copy l lives entirely in
//    the l-th group of four banks and is read by the lanes with (lane & 7) == l, so the eight lanes
//    that share a wavefront of a 128-bit load never meet in a bank whatever their intervals are: two
//    conflict-free LDS.128 per (row, term).  6 KB per term; with too many terms for that, rep = 1
//    (template parameter REP: the second half of an entry sits at a compile-time offset).
//    floor(s) is taken with a round-down add of 2^52 (no F2I / I2F conversions).  Rows outside the
//    edge knots and every other kind of term go through eval_term as before.
//  * one CTA walks tiles of R rows (round robin over a persistent grid).  Thread 0 issues one TMA bulk
//    copy per input column of the tile NS-1 steps ahead (completion on the stage's `full` mbarrier);
//    a warp that has read its rows of a stage arrives on the stage's `empty` mbarrier, which is all
//    thread 0 waits for before it refills the stage: no CTA-wide barrier inside the loop.
//  * the link / distribution switches are resolved at compile time for the canonical families (FAMK).
//  * rows after the last full tile, and inputs that are not 16-byte aligned, are read with plain loads.
#define PGB_PASS_MAX_FAST 48
struct PassSpl {
  double lo, scale, dn;  // s = (x - lo) * scale is the position in knot intervals, scale = nint / span
  int32_t feature, nint, coef, toff;  // toff: byte offset of the feature's column in a staged tile
};
#define PGB_PASS_MAX_TE 8
struct PassTe {  // te() of two cubic marginals: s_q = (x_q - lo_q) * scale_q, nint_q intervals, k1 = n_splines of marginal 1
  double lo0, scale0, lo1, scale1;
  int32_t f0, f1, n0, n1, k1, coef;
};
__host__ __device__ inline bool pass_te_eligible(const DTerm& T) {
  return T.kind == PGB_TERM_TENSOR && T.n_marg == 2 && T.by < 0 && T.marg[0].order == 3 && T.marg[1].order == 3 &&
         !T.marg[0].periodic && !T.marg[1].periodic;
}
struct PassPlan {
  int32_t n_fast, n_other, icpt, tab_doubles, fast_limit, pitch;  // pitch: table entries per term (max nint + 2)
  int32_t n_te, te_limit, rep_shift, pad;                          // 1 << rep_shift copies of every table
  PassTe te[PGB_PASS_MAX_TE];
  int16_t term[PGB_PASS_MAX_FAST];
  PassSpl spl[PGB_PASS_MAX_FAST];
};

static void pass_plan_build(const pgb_model* M, PassPlan& P, int rep_shift) {
  P.n_fast = P.n_other = P.tab_doubles = P.pitch = P.n_te = 0;
  P.rep_shift = rep_shift;
  P.pad = 0;
  P.te_limit = M->n_terms;
  P.icpt = -1;
  P.fast_limit = M->n_terms;
  for (int q = 0; q < M->n_terms; ++q) {
    const DTerm& T = M->dterms[q];
    if (T.kind == PGB_TERM_INTERCEPT && P.icpt < 0) { P.icpt = T.coef_offset; continue; }
    if (T.fast && P.n_fast < PGB_PASS_MAX_FAST && q < P.fast_limit) {
      const DMarg& g = T.marg[0];
      P.term[P.n_fast] = (int16_t)q;
      P.spl[P.n_fast++] = PassSpl{g.lo, g.inv_span * (double)g.nint, (double)g.nint, g.feature, g.nint, T.coef_offset, 0};
      P.pitch = std::max(P.pitch, g.nint + 2);
      continue;
    }
    if (pass_te_eligible(T) && q < P.te_limit) {
      if (P.n_te < PGB_PASS_MAX_TE) {
        const DMarg& a = T.marg[0];
        const DMarg& b = T.marg[1];
        P.te[P.n_te++] = PassTe{a.lo, a.inv_span * (double)a.nint, b.lo, b.inv_span * (double)b.nint,
                                a.feature, b.feature, a.nint, b.nint, b.nspl, T.coef_offset};
        continue;
      }
      P.te_limit = q;  // tensor terms from here on take the general path
    }
    if (T.fast && P.fast_limit == M->n_terms) P.fast_limit = q;  // fast terms from here on take the general path
    ++P.n_other;
  }
  P.tab_doubles = (4 * P.pitch * P.n_fast) << rep_shift;
}

// the terms that are not in the plan (f(), l(), te(), by=, other orders, cyclic bases): the general evaluation
template <class Load>
__device__ __forceinline__ double pass_other_terms(const DTerm* s_terms, int n_terms, int icpt, int fast_limit, int te_limit,
                                                   const double* s_beta, const Load& load) {
  EmitLp e{s_beta, 0.0};
  for (int q = 0; q < n_terms; ++q) {
    const DTerm& T = s_terms[q];
    if ((T.kind == PGB_TERM_INTERCEPT && T.coef_offset == icpt) || (T.fast && q < fast_limit)) continue;
    if (q < te_limit && pass_te_eligible(T)) continue;
    eval_term(T, load, e);
  }
  return e.lp;
}

// One cubic marginal of a planned te() term at s = position in knot intervals.  EXT = false assumes 0 <= s < nint and
// returns a non-zero flag when that fails (the interval index is clamped for memory safety only); EXT = true is the
// redo of a flagged row: the tangent of the basis at the edge knots outside them (utils.py:744-763), NaN for a NaN.
template <bool EXT>
__device__ __forceinline__ unsigned pass_te_marg(double s, int ni, int& i, double (&v)[4]) {
  if (EXT) {
    if (s < 0.0 || s > (double)ni) {
      const bool below = s < 0.0;
      const double h = 0.5 * (below ? s : s - (double)ni);
      i = below ? 0 : ni - 1;
      v[0] = below ? 1.0 / 6.0 - h : 0.0;
      v[1] = below ? 2.0 / 3.0 : 1.0 / 6.0 - h;
      v[2] = below ? 1.0 / 6.0 + h : 2.0 / 3.0;
      v[3] = below ? 0.0 : 1.0 / 6.0 + h;
      return 0u;
    }
    if (!(s >= 0.0)) {  // NaN
      i = 0;
      v[0] = v[1] = v[2] = v[3] = s;
      return 0u;
    }
    i = min((int)s, ni - 1);
    cubic_bspline(s - (double)i, v[0], v[1], v[2], v[3]);
    return 0u;
  }
  const double r = __dadd_rd(s, 4503599627370496.0);  // 2^52 + floor(s), exact for 0 <= s < 2^32
  const unsigned iu = (unsigned)__double2loint(r);
  i = (int)min(iu, (unsigned)(ni - 1));
  cubic_bspline(s - (r - 4503599627370496.0), v[0], v[1], v[2], v[3]);
  return (unsigned)(__double2hiint(r) ^ 0x43300000) | (unsigned)(iu >= (unsigned)ni);
}

template <bool EXT, class Load>
__device__ __forceinline__ double pass_te_term(const PassTe& T, const double* s_beta, const Load& load, unsigned& bad) {
  int i0, i1;
  double a[4], b[4];
  bad |= pass_te_marg<EXT>((load(T.f0) - T.lo0) * T.scale0, T.n0, i0, a);
  bad |= pass_te_marg<EXT>((load(T.f1) - T.lo1) * T.scale1, T.n1, i1, b);
  const double* cb = s_beta + T.coef + i0 * T.k1 + i1;  // first marginal slowest (utils.py:943-946)
  double v = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double* c = cb + r * T.k1;
    v = fma(a[r], fma(b[3], c[3], fma(b[2], c[2], fma(b[1], c[1], b[0] * c[0]))), v);
  }
  return v;
}

// lp of NR rows (row k through load[k]) -- the fast terms of all NR rows side by side (independent dependency chains).
// Table of term q, as 16-byte pairs (c_2h, c_2h+1), h = 0, 1, of entry e: REP = 8: pair[2 q pitch 8 + (2 e + h) 8 + l],
// l = copy; REP = 1: pair[(2 q + h) pitch + e].  e = 1 + floor(s) clamped to 0 .. nint + 1.  Entry e in
// 1 .. nint is the cubic of knot interval e - 1; entries 0 and nint + 1 are the tangent lines at the two edge knots,
// which is exactly the reference's linear extrapolation (utils.py:744-763); all in powers of s - (e - 1).
// The first loop assumes 0 <= s < nint and only notes when that fails (a feature outside the edge knots or on the last
// one, or a NaN); such a row is evaluated again by the second loop, which clamps the entry (a NaN stays a NaN: fmax
// drops it from the entry choice, the power keeps it).
template <int REP>
__device__ __forceinline__ int pass_tab_index(int e, int h, int l, int pitch) {
  return REP == 8 ? (2 * e + h) * 8 + l : h * pitch + e;
}

template <int NR, bool OTHER, int REP, class Load>
__device__ __forceinline__ void pass_lp(const PassPlan& P, const DTerm* s_terms, int n_terms, const double* s_beta,
                                        const double* s_tab, double icpt, const Load (&load)[NR], double (&lp)[NR]) {
  const double kC = 4503599627370497.0;  // 2^52 + 1: RD(s + kC) = kC + floor(s), exact for -1 <= s < 2^32
  bool bad[NR];
#pragma unroll
  for (int k = 0; k < NR; ++k) { lp[k] = icpt; bad[k] = false; }
  const int pitch = P.pitch;
  const int hi_off = REP == 8 ? 8 : pitch;  // pairs between the two halves of an entry
  const int stride = 2 * pitch * REP;       // pairs per term
  const double2* tab0 = reinterpret_cast<const double2*>(s_tab) + (REP == 8 ? (threadIdx.x & 7) : 0);
  {
    const double2* tab = tab0;
    for (int q = 0; q < P.n_fast; ++q, tab += stride) {
      const PassSpl& S = P.spl[q];
      const double lo = S.lo, scale = S.scale;
      const int f = S.feature, toff = S.toff;
      const unsigned ni = (unsigned)S.nint;
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const double s = (load[k].at(f, toff) - lo) * scale;
        const double r = __dadd_rd(s, kC);
        const unsigned e = (unsigned)__double2loint(r);
        bad[k] = bad[k] || (e - 1u >= ni) || (__double2hiint(r) != 0x43300000);  // ISETP.OR into one predicate
        const double2* tb = tab + (REP == 8 ? min(e, ni + 1u) << 4 : min(e, ni + 1u));  // in bounds whatever s was
        const double t = s - (r - kC);
        const double2 lo2 = tb[0], hi2 = tb[hi_off];
        lp[k] += fma(fma(fma(hi2.y, t, hi2.x), t, lo2.y), t, lo2.x);
      }
    }
  }
  for (int q = 0; q < P.n_te; ++q) {
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      unsigned b = 0u;
      lp[k] += pass_te_term<false>(P.te[q], s_beta, load[k], b);
      bad[k] = bad[k] || (b != 0u);
    }
  }
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    if (bad[k]) {
      double v = icpt;
      const double2* tab = tab0;
#pragma unroll 1
      for (int q = 0; q < P.n_fast; ++q, tab += stride) {
        const PassSpl& S = P.spl[q];
        const double s = (load[k](S.feature) - S.lo) * S.scale;
        const double r = __dadd_rd(fmin(fmax(s, -1.0), S.dn), kC);
        const double2* tb = tab + (REP == 8 ? __double2loint(r) << 4 : __double2loint(r));
        const double t = s - (r - kC);
        const double2 lo2 = tb[0], hi2 = tb[hi_off];
        v += fma(fma(fma(hi2.y, t, hi2.x), t, lo2.y), t, lo2.x);
      }
      unsigned b = 0u;
#pragma unroll 1
      for (int q = 0; q < P.n_te; ++q) v += pass_te_term<true>(P.te[q], s_beta, load[k], b);
      lp[k] = v;
    }
  }
  if (OTHER) {
#pragma unroll
    for (int k = 0; k < NR; ++k) lp[k] += pass_other_terms(s_terms, n_terms, P.icpt, P.fast_limit, P.te_limit, s_beta, load[k]);
  }
}

// What a pass does with the linear predictor of a row (template parameter FIN)
enum { PASS_STATS = 0,   // mu, deviance and the sums of pgb_stats_pass
       PASS_FULL = 1,    // + log-likelihood and the null model (scale known)
       PASS_IRLS = 2 };  // the IRLS step of a Gram pass: omega, omega z (written to lp_out, mu_out) and the PGB_GS_* sums

// link, deviance and the statistics sums of one row; wi: the weight, for PASS_IRLS its float32 reciprocal (weight_inv)
template <bool HAS_Y, int FIN>
__device__ __forceinline__ void pass_finish(const Family& fam, double lp, int64_t gi, double yi, double wi, double scale,
                                            double ybar, int poisson_round, double* __restrict__ lp_out,
                                            double* __restrict__ mu_out, double (&acc)[PGB_SS_COUNT]) {
  if (FIN == PASS_IRLS) {  // the same row arithmetic and sums as gram_prepass_kernel (pgb_cells.cu)
    const IrlsRow ir = irls_row(fam, lp, yi, wi);
    const double o = ir.omega, oz = ir.omega * ir.z;
    if (ir.keep) {
      acc[PGB_GS_NUSED] += 1.0;
      acc[PGB_GS_DEV] += dist_dev(fam, yi, ir.mu);
      acc[PGB_GS_NCORRECT] += (yi == ((ir.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
    }
    acc[PGB_GS_NROWS] += 1.0;
    acc[PGB_GS_COUNT] += o;
    acc[PGB_GS_COUNT + 1] += oz;
    lp_out[gi] = o;
    mu_out[gi] = oz;
    return;
  }
  const double mu = link_inv(fam, lp);
  if (lp_out) lp_out[gi] = lp;
  if (mu_out) mu_out[gi] = mu;
  if (!isfinite(lp) || !isfinite(mu)) acc[PGB_SS_NONFINITE] += 1.0;
  if (HAS_Y) {
    const double d = dist_dev(fam, yi, mu);
    const double res = yi - mu;
    acc[PGB_SS_N] += 1.0;
    acc[PGB_SS_WDEV] += wi * d;
    acc[PGB_SS_PEARSON] += wi * (res * res) / dist_var(fam, mu);  // distributions.py:78
    acc[PGB_SS_DEV] += d;
    acc[PGB_SS_NCORRECT] += (yi == ((mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
    acc[PGB_SS_SUMY] += yi;
    acc[PGB_SS_SUMW] += wi;
    if (FIN == PASS_FULL) {
      const double yl = poisson_round ? rint(yi * wi) : yi;  // pygam.py:2797-2798
      acc[PGB_SS_LOGLIK] += dist_logpdf(fam, yl, mu, wi, scale);
      acc[PGB_SS_NULL_WDEV] += wi * dist_dev(fam, yi, ybar);
      acc[PGB_SS_NULL_LOGLIK] += dist_logpdf(fam, yl, ybar, wi, scale);
    }
  }
}

template <int FIN>
__device__ __forceinline__ double pass_weight(const float* __restrict__ w, int64_t i) {
  if (FIN == PASS_IRLS) return weight_inv(w, i);
  return w ? (double)__ldg(w + i) : 1.0;
}

// kPassThreads consumer threads (RPT rows each per tile) + one producer warp
template <bool HAS_Y, int FIN, int FAMK, int RPT, bool OTHER, int REP>
__global__ void __launch_bounds__(kPassThreads + 32, 1)
stats_pass_kernel(const __grid_constant__ PassPlan P, const DTerm* __restrict__ g_terms, int n_terms, int m, int F,
                  Family fam_in, const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta, double scale, double ybar,
                  int poisson_round, int NS, int64_t n_bulk_tiles, double* __restrict__ part,
                  double* __restrict__ lp_out, double* __restrict__ mu_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int R = RPT * kPassThreads;
  Family fam = fam_in;
  if (FAMK == 1) { fam.dist = PGB_DIST_NORMAL; fam.link = PGB_LINK_IDENTITY; }
  if (FAMK == 2) { fam.dist = PGB_DIST_BINOMIAL; fam.link = PGB_LINK_LOGIT; }
  if (FAMK == 3) { fam.dist = PGB_DIST_POISSON; fam.link = PGB_LINK_LOG; }
  const int ncol = F + (HAS_Y ? 1 : 0);
  double* tiles = reinterpret_cast<double*>(smem_raw);                      // [NS][ncol][R]
  float* wtiles = reinterpret_cast<float*>(tiles + (size_t)NS * ncol * R);  // [NS][R]
  double* s_beta = reinterpret_cast<double*>(wtiles + (size_t)NS * R);
  double* s_tab = s_beta + ((m + 1) & ~1);
  double* s_red = s_tab + P.tab_doubles;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(s_red + PGB_SS_COUNT * 32);  // full[NS], empty[NS]
  DTerm* s_terms = reinterpret_cast<DTerm*>(mbar + 2 * NS);
  const int tid = threadIdx.x;
  for (int i = tid; i < m; i += blockDim.x) s_beta[i] = beta[i];
  load_terms_smem(s_terms, g_terms, n_terms);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_addr(&mbar[s]), 1);
      mbar_init(smem_addr(&mbar[NS + s]), kPassThreads / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // one cubic per (fast term, knot interval): p(t) = sum_a beta[c + a] b_a(t) in powers of t; two tangent lines
  for (int q = 0; q < P.n_fast; ++q) {
    const PassSpl& S = P.spl[q];
    for (int el = tid; el < (S.nint + 2) * REP; el += blockDim.x) {
      const int e = el / REP, l = el % REP;
      const int ci = e == 0 ? 0 : (e > S.nint ? S.nint - 1 : e - 1);
      const double* b = s_beta + S.coef + ci;
      double c0 = (b[0] + 4.0 * b[1] + b[2]) * (1.0 / 6.0);
      double c1 = (b[2] - b[0]) * 0.5;
      double c2 = (b[0] - 2.0 * b[1] + b[2]) * 0.5;
      double c3 = ((b[3] - b[0]) + 3.0 * (b[1] - b[2])) * (1.0 / 6.0);
      if (e == 0) { c0 -= c1; c2 = 0.0; c3 = 0.0; }  // below the first knot: p(0) + p'(0) s, in powers of s + 1
      if (e > S.nint) {                              // above the last knot: p(1) + p'(1) (s - nint)
        c0 = (b[1] + 4.0 * b[2] + b[3]) * (1.0 / 6.0);
        c1 = (b[3] - b[1]) * 0.5;
        c2 = 0.0; c3 = 0.0;
      }
      double2* tb = reinterpret_cast<double2*>(s_tab) + (size_t)2 * q * P.pitch * REP;
      tb[pass_tab_index<REP>(e, 0, l, P.pitch)] = make_double2(c0, c1);
      tb[pass_tab_index<REP>(e, 1, l, P.pitch)] = make_double2(c2, c3);
    }
  }
  __syncthreads();
  const bool use_w = HAS_Y && (w != nullptr);
  const size_t stage_doubles = (size_t)ncol * R;

  if (tid >= kPassThreads) {
    // ---- producer warp: one lane keeps NS tiles in flight ----
    if (tid == kPassThreads) {
      int stage = 0;
      int64_t j = 0;
      for (int64_t t = blockIdx.x; t < n_bulk_tiles; t += gridDim.x, ++j) {
        if (j >= NS) mbar_wait(smem_addr(&mbar[NS + stage]), (uint32_t)(((j / NS) - 1) & 1));
        const int64_t row0 = t * R;
        double* dst = tiles + (size_t)stage * stage_doubles;
        const uint32_t bar = smem_addr(&mbar[stage]);
        mbar_expect_tx(bar, (uint32_t)(ncol * R * 8 + (use_w ? R * 4 : 0)));
        for (int f = 0; f < F; ++f) bulk_g2s(smem_addr(dst + (size_t)f * R), X + (int64_t)f * ld + row0, R * 8, bar);
        if (HAS_Y) bulk_g2s(smem_addr(dst + (size_t)F * R), y + row0, R * 8, bar);
        if (use_w) bulk_g2s(smem_addr(wtiles + (size_t)stage * R), w + row0, R * 4, bar);
        if (++stage == NS) stage = 0;
      }
    }
    return;
  }

  // ---- consumers ----
  const double icpt = P.icpt >= 0 ? s_beta[P.icpt] : 0.0;
  double acc[PGB_SS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_SS_COUNT; ++s) acc[s] = 0.0;
  // full tiles from the ring.  The predict instances with one row per thread take two tiles before they finish either:
  // the inverse link of a row is one dependency chain (exp, a division), and two of them side by side give the warp more
  // instructions to issue (config 2: 65 -> 68 % of the HBM peak; the statistics instances did not gain, measured).
  {
    int stage = 0;
    uint32_t round = 0;
    constexpr int NT = (RPT == 1 && !HAS_Y) ? 2 : 1;  // tiles per trip
    for (int64_t t = blockIdx.x; t < n_bulk_tiles; t += (int64_t)NT * gridDim.x) {
      double lp[NT][RPT], yv[NT][RPT], wv[NT][RPT];
      const bool two = NT == 2 && t + gridDim.x < n_bulk_tiles;
#pragma unroll
      for (int u = 0; u < NT; ++u) {
        if (u == 1 && !two) break;
        mbar_wait(smem_addr(&mbar[stage]), round & 1u);
        const double* tile = tiles + (size_t)stage * stage_doubles;
        const float* wt = wtiles + (size_t)stage * R;
        LoadTile load[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) load[k] = LoadTile{tile, R, tid + k * kPassThreads};
        pass_lp<RPT, OTHER, REP>(P, s_terms, n_terms, s_beta, s_tab, icpt, load, lp[u]);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = tid + k * kPassThreads;
          yv[u][k] = HAS_Y ? tile[(size_t)F * R + r] : 0.0;
          wv[u][k] = !use_w ? 1.0 : FIN == PASS_IRLS ? (double)__fdiv_rn(1.0f, wt[r]) : (double)wt[r];
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_addr(&mbar[NS + stage]));  // the stage can be refilled
        if (++stage == NS) { stage = 0; ++round; }
      }
      const int64_t row0 = t * R + tid;
      if (two) {
#pragma unroll
        for (int u = 0; u < NT; ++u)
          pass_finish<HAS_Y, FIN>(fam, lp[u][0], row0 + (int64_t)u * gridDim.x * R, yv[u][0], wv[u][0], scale, ybar, poisson_round, lp_out, mu_out, acc);
      } else {
#pragma unroll
        for (int k = 0; k < RPT; ++k)
          pass_finish<HAS_Y, FIN>(fam, lp[0][k], row0 + k * kPassThreads, yv[0][k], wv[0][k], scale, ybar, poisson_round, lp_out, mu_out, acc);
      }
    }
  }
  // the remaining rows (and inputs that are not 16-byte aligned) with plain loads
  const int64_t step = (int64_t)gridDim.x * kPassThreads;
  int64_t i = n_bulk_tiles * R + (int64_t)blockIdx.x * kPassThreads + tid;
  if (RPT == 2) {  // two rows per thread here as well: twice the loads in flight when nothing is staged
    for (; i + step < n; i += 2 * step) {
      const LoadGlobal load[2] = {LoadGlobal{X, ld, i}, LoadGlobal{X, ld, i + step}};
      double lp[2];
      pass_lp<2, OTHER, REP>(P, s_terms, n_terms, s_beta, s_tab, icpt, load, lp);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int64_t ik = i + k * step;
        pass_finish<HAS_Y, FIN>(fam, lp[k], ik, HAS_Y ? __ldg(y + ik) : 0.0, pass_weight<FIN>(use_w ? w : nullptr, ik), scale, ybar,
                                 poisson_round, lp_out, mu_out, acc);
      }
    }
  }
  for (; i < n; i += step) {
    const LoadGlobal load[1] = {LoadGlobal{X, ld, i}};
    double lp[1];
    pass_lp<1, OTHER, REP>(P, s_terms, n_terms, s_beta, s_tab, icpt, load, lp);
    pass_finish<HAS_Y, FIN>(fam, lp[0], i, HAS_Y ? __ldg(y + i) : 0.0, pass_weight<FIN>(use_w ? w : nullptr, i), scale, ybar, poisson_round,
                            lp_out, mu_out, acc);
  }
  // fixed-order sum of the consumer threads (the producer warp has left: a named barrier of the consumers only)
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int s = 0; s < PGB_SS_COUNT; ++s) {
      double v = acc[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[s * 32 + warp] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kPassThreads) : "memory");
    constexpr int NPART = FIN == PASS_IRLS ? PGB_GS_COUNT + 2 : PGB_SS_COUNT;  // the Gram pass reduces PGB_GS_COUNT + 2 sums per CTA
    if (tid < NPART) {
      double v = 0.0;
      for (int wq = 0; wq < kPassThreads / 32; ++wq) v += s_red[tid * 32 + wq];
      part[(int64_t)blockIdx.x * NPART + tid] = v;
    }
  }
}

template <bool HAS_Y, int FIN, int FAMK, int RPT, bool OTHER, int REP>
static int stats_pass_launch(const PassPlan& P, const pgb_model* M, int grid, size_t smem, cudaStream_t st, const double* X,
                             int64_t n, int64_t ld, const double* y, const float* w, const double* beta, double scale,
                             double ybar, int poisson_round, int NS, int64_t n_bulk_tiles, double* part, double* lp_out,
                             double* mu_out) {
  PGB_CUDA(cudaFuncSetAttribute(stats_pass_kernel<HAS_Y, FIN, FAMK, RPT, OTHER, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kSmemMax));
  stats_pass_kernel<HAS_Y, FIN, FAMK, RPT, OTHER, REP><<<grid, kPassThreads + 32, smem, st>>>(
      P, M->d_terms, M->n_terms, M->m, M->n_features, M->fam, X, n, ld, y, w, beta, scale, ybar, poisson_round, NS,
      n_bulk_tiles, part, lp_out, mu_out);
  return PGB_OK;
}

// One O(k)-per-row pass over the rows: plan, shared-memory budget, grid and the template instance.  fin: PASS_STATS /
// PASS_FULL / PASS_IRLS; part receives the per-CTA sums (*grid_out CTAs).
static int pass_run(const pgb_model* M, const double* X, int64_t n, int64_t ld, const double* y, const float* w,
                    const double* beta, double scale, double ybar, int32_t poisson_round, int fin, double* part,
                    double* lp_out, double* mu_out, int* grid_out, cudaStream_t st) {
  const int F = M->n_features, ncol = F + (y ? 1 : 0);
  const size_t rowbytes = (size_t)ncol * 8 + 4;
  const size_t budget = kSmemMax - 1024;
  // one CTA of kPassThreads consumers per SM; stages of R = RPT * kPassThreads rows.  Eight copies of the interval
  // tables when that leaves three stages, one copy otherwise; two rows per thread when four stages fit.
  PassPlan P;
  size_t fixed = 0;
  int NS = 0, RPT = 1;
  for (int rep_shift = 3; rep_shift >= 0; rep_shift -= 3) {
    pass_plan_build(M, P, rep_shift);
    fixed = (size_t)((M->m + 1) & ~1) * 8 + (size_t)P.tab_doubles * 8 + PGB_SS_COUNT * 32 * 8 + 2 * 8 * 8 +
            (size_t)M->n_terms * sizeof(DTerm);
    const size_t stages1 = budget > fixed ? (budget - fixed) / (kPassThreads * rowbytes) : 0;
    if (stages1 >= 3 || rep_shift == 0) {
      if (stages1 >= 2) {
        if (stages1 >= 4 && fin != PASS_FULL && !P.n_other) RPT = 2;
        NS = (int)std::min<size_t>(4, stages1 / RPT);
      }
      break;
    }
  }
  if (NS < 2) {  // rows too wide for two stages (F > ~50): no staging, every row through the plain-load path
    NS = 0;
    RPT = 1;
    if (fixed > budget) {
      pgb_set_error("pgb_stats_pass: the model (%d coefficients) does not fit shared memory", M->m);
      return PGB_EUNSUPPORTED;
    }
  }
  const int R = RPT * kPassThreads;
  for (int q = 0; q < P.n_fast; ++q) P.spl[q].toff = P.spl[q].feature * R * 8;
  const size_t smem = fixed + (size_t)NS * R * rowbytes;
  // TMA bulk copies need 16-byte aligned sources: every column start and every tile start
  const bool bulk_ok = ((uintptr_t)X % 16 == 0) && (ld % 2 == 0) && (!y || (uintptr_t)y % 16 == 0) &&
                       (!w || (uintptr_t)w % 16 == 0);
  const int64_t n_bulk_tiles = (bulk_ok && NS >= 2) ? n / R : 0;
  const int64_t n_units = std::max<int64_t>(n_bulk_tiles, (n - n_bulk_tiles * R + kPassThreads - 1) / kPassThreads);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_units, std::min(M->pass_ctas, kSMs)));
  const Family& fm = M->fam;
  const int famk = (fm.dist == PGB_DIST_NORMAL && fm.link == PGB_LINK_IDENTITY)   ? 1
                   : (fm.dist == PGB_DIST_BINOMIAL && fm.link == PGB_LINK_LOGIT) ? 2
                   : (fm.dist == PGB_DIST_POISSON && fm.link == PGB_LINK_LOG)    ? 3
                                                                                 : 0;
  int rc = PGB_OK;
#define PGB_STATS_ARGS P, M, grid, smem, st, X, n, ld, y, w, beta, scale, ybar, poisson_round, NS, n_bulk_tiles, part, lp_out, mu_out
#define PGB_STATS_REP(HY, FL, FK, RP, OT)                                        \
  rc = P.rep_shift ? stats_pass_launch<HY, FL, FK, RP, OT, 8>(PGB_STATS_ARGS)       \
                   : stats_pass_launch<HY, FL, FK, RP, OT, 1>(PGB_STATS_ARGS);
#define PGB_STATS_FAM(HY, FL, RP, OT)                        \
  switch (famk) {                                            \
    case 1: PGB_STATS_REP(HY, FL, 1, RP, OT) break;          \
    case 2: PGB_STATS_REP(HY, FL, 2, RP, OT) break;          \
    case 3: PGB_STATS_REP(HY, FL, 3, RP, OT) break;          \
    default: PGB_STATS_REP(HY, FL, 0, RP, OT) break;         \
  }
#define PGB_STATS_RPT(HY, FL)                                  \
  if (P.n_other) { PGB_STATS_FAM(HY, FL, 1, true) }            \
  else if (RPT == 2) { PGB_STATS_FAM(HY, FL, 2, false) }       \
  else { PGB_STATS_FAM(HY, FL, 1, false) }
  if (fin == PASS_FULL) {
    if (P.n_other) { PGB_STATS_FAM(true, PASS_FULL, 1, true) }
    else { PGB_STATS_FAM(true, PASS_FULL, 1, false) }
  } else if (fin == PASS_IRLS) { PGB_STATS_RPT(true, PASS_IRLS) }
  else if (!y) { PGB_STATS_RPT(false, PASS_STATS) }
  else { PGB_STATS_RPT(true, PASS_STATS) }
#undef PGB_STATS_RPT
#undef PGB_STATS_FAM
#undef PGB_STATS_REP
#undef PGB_STATS_ARGS
  if (rc != PGB_OK) return rc;
  PGB_LAUNCHED(1);
  PGB_CUDA(cudaGetLastError());
  *grid_out = grid;
  return PGB_OK;
}

extern "C" int pgb_stats_pass(const pgb_model* M, const double* X, int64_t n, int64_t ld,
                              const double* y, const float* w, const double* beta, double scale,
                              double ybar, int32_t poisson_round, double* scalars, double* lp_out,
                              double* mu_out, void* ws, int64_t ws_bytes, void* stream) {
  if (!M || !X || !beta || n < 1 || ld < n) { pgb_set_error("pgb_stats_pass: bad arguments"); return PGB_EINVAL; }
  if (ws_bytes < pgb_stats_ws_bytes(M) || !ws) { pgb_set_error("pgb_stats_pass: workspace too small"); return PGB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0;
  const int rc = pass_run(M, X, n, ld, y, w, beta, scale, ybar, poisson_round, (y && scale > 0.0) ? PASS_FULL : PASS_STATS,
                          (double*)ws, lp_out, mu_out, &grid, st);
  if (rc != PGB_OK) return rc;
  if (scalars) {
    scalar_reduce_kernel<<<1, 32 * PGB_SS_COUNT, 0, st>>>((const double*)ws, grid, PGB_SS_COUNT, scalars, 0);
    PGB_LAUNCHED(1);
    PGB_CUDA(cudaGetLastError());
  }
  return PGB_OK;
}

// The IRLS step of a Gram pass over a prepared shard (pgb_cells.cu): lp from X through the pass above, omega and
// omega z per row, the PGB_GS_* sums and sum omega, sum omega z per CTA in scal_part (*grid_out parts).  Nothing the
// prepared shard stores is read or written.
int pgb_pass_irls(const pgb_model* M, const double* X, int64_t n, int64_t ld, const double* y, const float* w,
                  const double* beta, double* om, double* omz, double* scal_part, int* grid_out, cudaStream_t st) {
  return pass_run(M, X, n, ld, y, w, beta, -1.0, 0.0, 0, PASS_IRLS, scal_part, om, omz, grid_out, st);
}
