/*
 * pgb.h -- C ABI of the B200-native PIRLS engine behind the pyGAM model API.
 *
 * pyGAM (the reference, /root/reference/pygam, v0.12.0) is pure Python and has no FFI of its
 * own; the seam this library sits behind is the private solver method
 * GAM._pirls (pygam/pygam.py:715-822) together with GAM._modelmat / _linear_predictor
 * (pygam.py:385-421, 469-497) and GAM._estimate_model_statistics (pygam.py:993-1057).  Each
 * entry point below names the reference code it replaces.  INTEGRATION.md shows the ctypes
 * binding a pyGAM maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success and a negative PGB_E* code otherwise; nothing throws
 *    across the ABI; pgb_last_error() returns a thread-local message for the last failure.
 *  - pointers named X, y, w, beta, out*, ws are DEVICE pointers unless the function name ends
 *    in _host.  The caller owns every buffer; the model handle owns only its descriptor tables.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it.
 *  - X is FEATURE-MAJOR: feature f of row i lives at X[f * ld + i], ld >= n (elements).
 *  - w (prior weights) is float32 or NULL: pyGAM casts user weights to float32
 *    (pygam.py:889); NULL means unit weights.
 *  - all arithmetic is IEEE fp64.
 */
#ifndef PGB_H_
#define PGB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_MAX_MARG 4   /* marginals of a tensor term (te) */
#define PGB_MAX_ORDER 7  /* spline_order upper bound      */

/* term kinds: pygam/terms.py Intercept 497-599, LinearTerm 602-708, SplineTerm 711-956,
 * FactorTerm 959-1101, TensorTerm 1104-1641 */
enum { PGB_TERM_INTERCEPT = 0, PGB_TERM_LINEAR = 1, PGB_TERM_SPLINE = 2, PGB_TERM_FACTOR = 3,
       PGB_TERM_TENSOR = 4 };

/* distributions: pygam/distributions.py 97-612 */
enum { PGB_DIST_NORMAL = 0, PGB_DIST_BINOMIAL = 1, PGB_DIST_POISSON = 2, PGB_DIST_GAMMA = 3,
       PGB_DIST_INVGAUSS = 4 };

/* links: pygam/links.py 21-314 */
enum { PGB_LINK_IDENTITY = 0, PGB_LINK_LOGIT = 1, PGB_LINK_LOG = 2, PGB_LINK_INVERSE = 3,
       PGB_LINK_INVSQUARED = 4 };

/* error codes */
enum { PGB_OK = 0, PGB_EINVAL = -1, PGB_ECUDA = -2, PGB_ENOMEM = -3, PGB_ENOTPD = -4,
       PGB_EDIVERGED = -5, PGB_EUNSUPPORTED = -6 };

/* One term of the model, flattened from Term.info (pygam/terms.py:220-233).  For a tensor term
 * the per-marginal arrays hold the marginal splines in order (first marginal varies slowest in
 * the coefficient index, pygam/utils.py:943-946). */
#define PGB_MARG_LINEAR (-1)
typedef struct pgb_term {
  int32_t kind;                     /* PGB_TERM_*                                           */
  int32_t n_marg;                   /* 1 for l/s/f, 2..PGB_MAX_MARG for te, 0 for intercept  */
  int32_t feature[PGB_MAX_MARG];    /* column of X                                           */
  int32_t n_splines[PGB_MAX_MARG];  /* basis functions of the marginal (f: number of levels) */
  int32_t order[PGB_MAX_MARG];      /* spline_order (f: 0); PGB_MARG_LINEAR for an l() marginal of te(): one column, the feature value */
  int32_t periodic[PGB_MAX_MARG];   /* basis == "cp"                                         */
  double edge_lo[PGB_MAX_MARG];     /* edge_knots_[0]                                        */
  double edge_hi[PGB_MAX_MARG];     /* edge_knots_[1]                                        */
  int32_t by;                       /* by-variable column or -1 (terms.py:953-954, 1474-75)  */
  int32_t drop_first;               /* f(coding="dummy"), terms.py:1093-1094                 */
  int32_t coef_offset;              /* first coefficient of the term                         */
  int32_t n_coefs;                  /* terms.py n_coefs                                      */
} pgb_term;

typedef struct pgb_model pgb_model;

/* sizes derived from the descriptor */
typedef struct pgb_model_info {
  int32_t m;            /* total coefficients                                   */
  int32_t k;            /* non-zeros per model-matrix row                       */
  int32_t n_features;   /* feature columns X must have                          */
  int32_t n_roles;      /* Gram accumulator partitions          */
  int32_t tile_rows;    /* rows per shared-memory tile of the Gram pass         */
  int32_t gram_ctas;    /* CTAs launched by the Gram pass                       */
  int64_t gram_ws_bytes;/* workspace pgb_gram_pass needs                        */
  int64_t pass_ws_bytes;/* workspace pgb_stats_pass needs                       */
  int32_t cell_roles;   /* > 0: the Gram pass runs cell-owned (all-spline model), this many roles */
  int32_t cell_tile_rows;/* rows per shared-memory tile of the cell-owned Gram pass */
  int32_t cell_threads; /* threads per CTA of the spline-pair kernel (256, 320, 384 or 512)   */
  int32_t cell_ring_groups; /* > 0: gram_cells_ring_kernel (one CTA per SM); 0: gram_cells_sorted_kernel, two CTAs per SM */
  int32_t cell_generic_roles; /* roles of the te()/f()/l() kernel among cell_roles           */
  int32_t reserved0;
} pgb_model_info;

const char* pgb_last_error(void);
int pgb_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long pgb_launch_count(void);

/* Build the device-side model descriptor.  Replaces the Python-side state that
 * TermList.compile (terms.py:1810-1833) leaves on the terms; `expectile` <= 0 disables the
 * asymmetric weight of ExpectileGAM._W (pygam.py:3449-3460).  device = CUDA ordinal. */
int pgb_model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist,
                     int32_t link, double levels, double expectile, int32_t device,
                     pgb_model** out);
void pgb_model_destroy(pgb_model* model);
/* TEST-ONLY entry point (tests/test_cells.py): the same model with the plan of its Gram pass constrained, so
 * that the parity tests can compare the two decompositions of the same sums (cell-owned register accumulation
 * vs the general shared-memory path) and exercise tile and band boundaries at small n.  The product never
 * calls it; all fields zero = pgb_model_create. */
typedef struct pgb_debug_opts {
  int32_t force_general_path; /* skip the cell-owned pass even where the model qualifies            */
  int32_t cells_tile_rows;    /* rows per shared-memory tile of the cell-owned pass (0: planned)    */
  int32_t cells_band_tiles;   /* tiles per L2 band (0: planned)                                     */
  int32_t cells_threads;      /* CTA width of the cell kernels: 256, 320, 384 or 512 (0: planned)   */
  int32_t cells_kernel;       /* sum kernel of a prepared shard: 1 barrier version, 2 ring (0: planned) */
} pgb_debug_opts;
int pgb_debug_model_create(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist,
                           int32_t link, double levels, double expectile, int32_t device,
                           const pgb_debug_opts* opts, pgb_model** out);
/* Host-only planning of the Gram pass (no CUDA call): fills info; *coverage_errors = entries of
 * the upper triangle of [G | r] not produced by exactly one accumulator element (0 = valid). */
int pgb_model_plan_host(const pgb_term* terms, int32_t n_terms, int32_t n_features,
                        pgb_model_info* info, int32_t* coverage_errors);
int pgb_model_get_info(const pgb_model* model, pgb_model_info* info);

/* gen_edge_knots (pygam/utils.py:592-621): per-feature min and max.  minmax: 2*F doubles
 * (device), [min_0, max_0, min_1, ...].  Also counts non-finite entries (check_X,
 * utils.py:180-182) into nonfinite[F] (device int64). */
int64_t pgb_col_stats_ws_bytes(int32_t F);
int pgb_col_stats(const double* X, int64_t n, int64_t ld, int32_t F, double* minmax,
                  long long* nonfinite, void* ws, int64_t ws_bytes, void* stream);

/* Distinct-level count of an integer-valued feature column (FactorTerm.compile,
 * terms.py:1071): marks present[(int)(x - lo)] for lo <= x <= lo + range-1; present has range+1
 * entries (device int32), the last one flags a non-integer value. */
int pgb_col_levels(const double* x, int64_t n, double lo, int32_t range, int32_t* present,
                   void* stream);

/* TermList.build_columns (terms.py:1901-1923) for inspection and small n: dense row-major
 * (n x m) model matrix. */
int pgb_model_matrix(const pgb_model* model, const double* X, int64_t n, int64_t ld,
                     double* out, void* stream);

/* scalars written by pgb_gram_pass (doubles) */
enum { PGB_GS_NUSED = 0,     /* rows kept by GAM._mask (pygam.py:638-663)                   */
       PGB_GS_DEV = 1,       /* unweighted, unscaled deviance on kept rows (callbacks.py:141) */
       PGB_GS_NCORRECT = 2,  /* #(y == (mu > 0.5)) on kept rows (callbacks.py:174)          */
       PGB_GS_ZWZ = 3,       /* sum of (W z)^2                                              */
       PGB_GS_NROWS = 4,     /* rows seen                                                   */
       PGB_GS_NONFINITE = 5, /* INIT mode: rows whose link(y') is not finite (pygam.py:706) */
       PGB_GS_COUNT = 8 };

enum { PGB_MODE_IRLS = 0,    /* weights and working response at beta (pygam.py:764-777)     */
       PGB_MODE_INIT = 1,    /* unweighted Gram and X^T link(y') of _initial_estimate
                                (pygam.py:696-710)                                          */
       PGB_MODE_PLANNED = 0x100 }; /* or-ed into mode: ws holds what pgb_gram_prepare left for this X */

/* The fused pass: evaluates the B-spline / tensor basis row in registers, the link and
 * distribution step, and accumulates G = X^T W^2 X, r = X^T W^2 z.  Replaces
 * pygam.py:764-783 (and b_spline_basis / build_columns, which are never materialised).
 * out: m*m + m + PGB_GS_COUNT doubles = [G (row-major, symmetric) | r | scalars], so that one
 * all-reduce of `out` combines row shards. */
int64_t pgb_gram_ws_bytes(const pgb_model* model, int64_t n);
/* smaller workspace for callers that never call pgb_gram_prepare (no PGB_MODE_PLANNED passes) */
int64_t pgb_gram_ws_bytes_streaming(const pgb_model* model, int64_t n);
int pgb_gram_pass(const pgb_model* model, int32_t mode, const double* X, int64_t n, int64_t ld,
                  const double* y, const float* w, const double* beta, double* out, void* ws,
                  int64_t ws_bytes, void* stream);
/* What does not change between the PIRLS iterations of one fit (pygam.py:741 builds the model matrix once,
 * before the loop): for an all-spline model the knot interval and local coordinate of every (row, term)
 * and, per term pair and row tile, the order of the rows by knot cell.  Writes them into `ws` (same
 * workspace and size as pgb_gram_pass); later calls of pgb_gram_pass for the same X, n and ws pass
 * mode | PGB_MODE_PLANNED and skip the basis search and the in-kernel sort.  No-op for other models. */
int pgb_gram_prepare(const pgb_model* model, const double* X, int64_t n, int64_t ld, void* ws,
                     int64_t ws_bytes, void* stream);

/* Optional second step of the preparation, for shards that will see many passes (long fits, a grid search over
 * the same rows): rewrites the stored order of the rows INSIDE every knot cell so that the gathers of the sum
 * kernels hit fewer shared-memory bank conflicts (cell sums of config 2: -13 % per pass; costs about three passes;
 * the sums of a cell are taken in another order, so [G | r] changes in the last bits).  ws: the workspace
 * pgb_gram_prepare filled for the same model and n.  No-op for models on the general path. */
int pgb_gram_refine(const pgb_model* model, int64_t n, void* ws, int64_t ws_bytes, void* stream);

/* Time the Gram pass with CUDA events: enable, run pgb_gram_pass, then read the device time (ms) of
 * the records kernel (basis + IRLS step -> row records) and of the accumulation kernel (records ->
 * G, r), each summed over the row chunks, and of the whole pass (with the reductions). */
int pgb_gram_timing(pgb_model* model, int32_t enable);
int pgb_gram_last_ms(pgb_model* model, float* records_ms, float* accum_ms, float* total_ms);

/* scalars written by pgb_stats_pass (doubles) */
enum { PGB_SS_N = 0, PGB_SS_WDEV = 1,  /* sum w * deviance(y, mu), unscaled (pygam.py:1215-1217) */
       PGB_SS_PEARSON = 2,             /* sum w (y-mu)^2 / V(mu) (distributions.py:78)          */
       PGB_SS_DEV = 3,                 /* unweighted deviance                                   */
       PGB_SS_NCORRECT = 4, PGB_SS_SUMY = 5, PGB_SS_SUMW = 6,
       PGB_SS_LOGLIK = 7,              /* sum log pdf(y | mu, scale) (distributions.py log_pdf) */
       PGB_SS_NULL_WDEV = 8,           /* sum w * deviance(y, ybar) (pygam.py:1141-1143)        */
       PGB_SS_NULL_LOGLIK = 9, PGB_SS_NONFINITE = 10, PGB_SS_COUNT = 16 };

/* lp = X beta, mu = link.mu(lp) and the deviance / likelihood sums of
 * _estimate_model_statistics (pygam.py:1031-1056) and predict_mu (pygam.py:423-450) on ALL
 * rows.  y, w, lp_out, mu_out may be NULL.  `scale` and `ybar` feed the log-likelihood and the
 * null-model sums; pass scale <= 0 to skip them.  poisson_round: PoissonGAM._loglikelihood
 * rounds y*w to integers (pygam.py:2797-2798). */
int64_t pgb_stats_ws_bytes(const pgb_model* model);
int pgb_stats_pass(const pgb_model* model, const double* X, int64_t n, int64_t ld,
                   const double* y, const float* w, const double* beta, double scale,
                   double ybar, int32_t poisson_round, double* scalars, double* lp_out,
                   double* mu_out, void* ws, int64_t ws_bytes, void* stream);

/* Linear predictor of one term (term >= 0) or of the whole model (term = -1) and its variance
 * var_i = x_i^T cov x_i per row, basis rows evaluated on the fly: GAM._get_quantiles
 * (pygam.py:1399-1404) over _modelmat / _linear_predictor restricted to a term (pygam.py:385-421,
 * 469-497) -- the n-sized part of confidence_intervals / prediction_intervals / partial_dependence
 * (pygam.py:1297-1334, 1520-1643, 2477-2506).  cov: m x m (device, symmetric); lp_out, var_out: n
 * doubles each, either may be NULL. */
int pgb_interval_pass(const pgb_model* model, const double* X, int64_t n, int64_t ld, const double* beta,
                      const double* cov, int32_t term, double* lp_out, double* var_out, void* stream);

/* ---- m x m penalised solve --------------------------------------------------------------
 * Replaces pygam.py:783-805 (QR of W.B, SVD of [R; E], back-projection) by the algebraically
 * equal beta = (G + E^T E)^-1 r, solved in a basis T = [Nhat | Z] that separates the analytic
 * null space of the design: T is given as `p` Householder reflectors
 * H_q = I - tau[q] v_q v_q^T (hv: p vectors of length m, one after the other; T = H_0 ... H_{p-1}).  EtE = S + P (+ C) in the ORIGINAL basis.  For a batch of
 * `nb` systems (grid-search candidates) EtE, EtE_null, out are strided by their own sizes while
 * G, r are shared when share_gram != 0.  The penalty is passed in two parts, EtE + EtE_null:
 * EtE_null (may be NULL) holds the components that annihilate the null vectors structurally
 * (difference penalties, constraint penalties D M D^T, which can be as large as 1e9); like G its
 * null block is set to exact zero after the change of basis, so that the sqrt(EPS)-sized
 * entries of the remaining part EtE are not swamped by its rounding noise.
 * out (per system, doubles): [beta (m) | edof | diff | info | logdet | ||beta|| | reserved(3)]
 * = m + 8.  beta_prev (m per system, or one shared vector when share_prev != 0; may be NULL)
 * feeds diff = ||beta_prev - beta|| / ||beta|| (pygam.py:804).  flags: PGB_SOLVE_EDOF computes
 * edof = tr(G (G+EtE)^-1) (pygam.py:1033-1034); PGB_SOLVE_COV also writes cov_out =
 * (G+EtE)^-1 G (G+EtE)^-1, m x m per system (pygam.py:1043-1045 without scale^2).
 * info: 0 ok, 1 = not positive definite, 2 = non-finite input (pygam.py:785-786). */
#define PGB_SOLVE_OUT_EXTRA 8
enum { PGB_SOLVE_EDOF = 1, PGB_SOLVE_COV = 2,
       PGB_SOLVE_PREPARED = 4 /* EtE holds the output of pgb_solve_prepare; EtE_null is ignored */ };
/* The penalty does not change during a PIRLS loop (pygam.py:746-805 builds S + P + C once per
 * iteration from the same matrices): its change of basis, A0 = Q^T EtE Q + [Q^T EtE_null Q]_0
 * (m x m per system), can be computed once and passed to pgb_solve as EtE with
 * PGB_SOLVE_PREPARED.  ws: nb * m * m doubles when EtE_null is given. */
int pgb_solve_prepare(int32_t m, int32_t nb, const double* EtE, const double* EtE_null, const double* hv,
                      const double* tau, int32_t p, double* A0, void* ws, int64_t ws_bytes, void* stream);
int64_t pgb_solve_ws_bytes(int32_t m, int32_t nb);
int pgb_solve(int32_t m, int32_t nb, const double* G, const double* r, int32_t share_gram,
              const double* EtE, const double* EtE_null, const double* hv, const double* tau, int32_t p,
              const double* beta_prev, int32_t share_prev, int32_t flags, double* out,
              double* cov_out, void* ws, int64_t ws_bytes, void* stream);

/* Rank-limited solve for fewer kept rows than coefficients: the reference crops its SVD of [R; E] to the leading
 * rank = min(m, n, rows kept by GAM._mask) singular triplets (pygam.py:789-799; pygam/tests/test_GAM_methods.py:93-107).
 * With A = G + EtE + EtE_null = V D^2 V^T (symmetric Jacobi eigen-decomposition on the device) this is
 * beta = V_k D_k^-2 V_k^T r over the k = rank largest eigenvalues, edof = tr(D_k^-1 V_k^T G V_k D_k^-1) and
 * cov = V_k D_k^-2 (V_k^T G V_k) D_k^-2 V_k^T.  EtE, EtE_null: the penalty in the ORIGINAL basis (EtE_null may be
 * NULL).  out, cov_out, flags (PGB_SOLVE_COV) as pgb_solve; ws: pgb_solve_truncated_ws_bytes(m). */
int64_t pgb_solve_truncated_ws_bytes(int32_t m);
int pgb_solve_truncated(int32_t m, int32_t rank, const double* G, const double* r, const double* EtE,
                        const double* EtE_null, const double* beta_prev, int32_t flags, double* out,
                        double* cov_out, void* ws, int64_t ws_bytes, void* stream);

/* Cholesky test of S + P (+ C) alone: GAM._cholesky (pygam.py:499-537, utils.py:30-91).
 * info_out (device int32): 0 = positive definite. */
int64_t pgb_chol_ws_bytes(int32_t m);
int pgb_chol_check(int32_t m, const double* A, int32_t* info_out, void* ws, int64_t ws_bytes,
                   void* stream);
/* the same test for nb matrices at once (the candidates of a grid search, pygam.py:2042-2070): A = nb x m x m,
 * info_out = nb int32, ws >= nb * m * m doubles (only used when a matrix does not fit shared memory) */
int pgb_chol_check_batch(int32_t m, int32_t nb, const double* A, int32_t* info_out, void* ws, int64_t ws_bytes,
                         void* stream);

/* ---- host-buffer entry points (end-to-end path) -----------------------------------------
 * Same as pgb_gram_pass but X, y, w live in (pinned) HOST memory, row-shard by row-shard:
 * chunks are copied to the device on two streams and overlapped with the Gram kernel.
 * out_host receives [G | r | scalars].  The staging buffers live in the model handle and are
 * reused by later calls. */
int pgb_gram_pass_host(pgb_model* model, int32_t mode, const double* X_host, int64_t n,
                       int64_t ld, const double* y_host, const float* w_host,
                       const double* beta_host, double* out_host, int64_t chunk_rows);

/* The rows of a fit from the caller's host array to the device shard: replaces the ingestion of GAM.fit / predict
 * (pygam.py:861-917 -> check_X / check_y, utils.py:96-263, which keep X as an (n, F) row-major host array).
 * src: (n, F) row-major doubles in pageable or pinned host memory; dst: device memory, column j of X goes to
 * dst[j * ld .. j * ld + n) (the feature-major layout every pass reads).  n_threads worker threads (0: half the
 * host cores, at most 8) transpose blocks of rows into pinned staging buffers owned by the library and send each block
 * with one strided copy on a stream of their own; the call returns when all rows are on the device.  dst must not
 * have work pending on another stream.  F = 1 uploads a vector (y). */
int pgb_upload_rows(double* dst, int64_t ld, const double* src, int64_t n, int32_t F, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* PGB_H_ */
