// pgb_pass.cu -- host side of the O(k)-per-row passes: pgb_stats_pass (statistics / predict, _linear_predictor
// pygam.py:385-421, predict_mu 423-450, the n-sized sums of _estimate_model_statistics 1031-1056) and the IRLS step of a
// Gram pass over a prepared shard (pgb_pass_irls, called by pgb_cells.cu).  The kernel is in pgb_pass.cuh.
#include "pgb_pass.cuh"

// instantiated in pgb_pass_b.cu
extern template int pass_dispatch<false, PASS_STATS>(const PassLaunch&);
extern template int pass_dispatch<true, PASS_IRLS>(const PassLaunch&);
template int pass_dispatch<true, PASS_STATS>(const PassLaunch&);
template int pass_dispatch<true, PASS_FULL>(const PassLaunch&);

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
  int NS = 0, RPT = 1, CT = kPassThreads;
  for (int rep_shift = 3; rep_shift >= 0; rep_shift -= 3) {
    pass_plan_build(M, P, rep_shift);
    fixed = (size_t)((M->m + 1) & ~1) * 8 + (size_t)P.tab_doubles * 8 + PGB_SS_COUNT * 32 * 8 + 2 * 8 * 8 +
            (size_t)M->n_terms * sizeof(DTerm);
    size_t stages1 = budget > fixed ? (budget - fixed) / (kPassThreads * rowbytes) : 0;
    if (stages1 >= 3 || rep_shift == 0) {
      // 576 consumers instead of 480 where three stages of the longer tiles still fit beside the eight table copies
      // (configs 2, 4, 5: +2 to +5 points of the HBM peak; config 3's 24 columns keep 480 and two stages)
      const size_t wide1 = budget > fixed ? (budget - fixed) / (kPassThreadsWide * rowbytes) : 0;
      if (rep_shift == 3 && !P.n_other && wide1 >= 3) {
        CT = kPassThreadsWide;
        stages1 = wide1;
      }
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
  const int R = RPT * CT;
  for (int q = 0; q < P.n_fast; ++q) P.spl[q].toff = P.spl[q].feature * R * 8;
  const size_t smem = fixed + (size_t)NS * R * rowbytes;
  // TMA bulk copies need 16-byte aligned sources: every column start and every tile start
  const bool bulk_ok = ((uintptr_t)X % 16 == 0) && (ld % 2 == 0) && (!y || (uintptr_t)y % 16 == 0) &&
                       (!w || (uintptr_t)w % 16 == 0);
  const int64_t n_bulk_tiles = (bulk_ok && NS >= 2) ? n / R : 0;
  const int64_t n_units = std::max<int64_t>(n_bulk_tiles, (n - n_bulk_tiles * R + CT - 1) / CT);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_units, std::min(M->pass_ctas, kSMs)));
  const Family& fm = M->fam;
  const int famk = (fm.dist == PGB_DIST_NORMAL && fm.link == PGB_LINK_IDENTITY)   ? 1
                   : (fm.dist == PGB_DIST_BINOMIAL && fm.link == PGB_LINK_LOGIT) ? 2
                   : (fm.dist == PGB_DIST_POISSON && fm.link == PGB_LINK_LOG)    ? 3
                                                                                 : 0;
  PassLaunch A{P, M, grid, NS, RPT, famk, CT, smem, st, X, n, ld, n_bulk_tiles, y, w, beta, scale, ybar, poisson_round, part, lp_out, mu_out};
  const int rc = fin == PASS_FULL   ? pass_dispatch<true, PASS_FULL>(A)
                 : fin == PASS_IRLS ? pass_dispatch<true, PASS_IRLS>(A)
                 : y                ? pass_dispatch<true, PASS_STATS>(A)
                                    : pass_dispatch<false, PASS_STATS>(A);
  if (rc != PGB_OK) return rc;
  pgb_count_launch(1);
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
    pgb_count_launch(1);
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
