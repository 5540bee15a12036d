// pgb_pass.cuh -- the O(k)-per-row pass kernel (statistics, predict, the IRLS step of a prepared Gram pass) and its
// template dispatch.  Included by pgb_pass.cu and pgb_pass_b.cu, which instantiate disjoint halves of the ~90 kernel
// instances so that they compile side by side.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "pgb_internal.h"
#include "pgb_kernels.cuh"

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

// CT consumer threads (RPT rows each per tile) + one producer warp.  CT = CT (480: 16 warps, 128 registers per
// thread) or CTWide (576: 19 warps, 104 registers -- the lean instances need 99) where the model leaves three
// stages of 576-row tiles: three more warps to hide the latency of the table gathers and of the link arithmetic
template <bool HAS_Y, int FIN, int FAMK, int RPT, bool OTHER, int REP, int CT>
__global__ void __launch_bounds__(CT + 32, 1)
stats_pass_kernel(const __grid_constant__ PassPlan P, const DTerm* __restrict__ g_terms, int n_terms, int m, int F,
                  Family fam_in, const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta, double scale, double ybar,
                  int poisson_round, int NS, int64_t n_bulk_tiles, double* __restrict__ part,
                  double* __restrict__ lp_out, double* __restrict__ mu_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int R = RPT * CT;
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
      mbar_init(smem_addr(&mbar[NS + s]), CT / 32);
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

  if (tid >= CT) {
    // ---- producer warp: one lane keeps NS tiles in flight ----
    if (tid == CT) {
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
        for (int k = 0; k < RPT; ++k) load[k] = LoadTile{tile, R, tid + k * CT};
        pass_lp<RPT, OTHER, REP>(P, s_terms, n_terms, s_beta, s_tab, icpt, load, lp[u]);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = tid + k * CT;
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
          pass_finish<HAS_Y, FIN>(fam, lp[0][k], row0 + k * CT, yv[0][k], wv[0][k], scale, ybar, poisson_round, lp_out, mu_out, acc);
      }
    }
  }
  // the remaining rows (and inputs that are not 16-byte aligned) with plain loads
  const int64_t step = (int64_t)gridDim.x * CT;
  int64_t i = n_bulk_tiles * R + (int64_t)blockIdx.x * CT + tid;
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
    asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
    constexpr int NPART = FIN == PASS_IRLS ? PGB_GS_COUNT + 2 : PGB_SS_COUNT;  // the Gram pass reduces PGB_GS_COUNT + 2 sums per CTA
    if (tid < NPART) {
      double v = 0.0;
      for (int wq = 0; wq < CT / 32; ++wq) v += s_red[tid * 32 + wq];
      part[(int64_t)blockIdx.x * NPART + tid] = v;
    }
  }
}

template <bool HAS_Y, int FIN, int FAMK, int RPT, bool OTHER, int REP, int CT>
static int stats_pass_launch(const PassPlan& P, const pgb_model* M, int grid, size_t smem, cudaStream_t st, const double* X,
                             int64_t n, int64_t ld, const double* y, const float* w, const double* beta, double scale,
                             double ybar, int poisson_round, int NS, int64_t n_bulk_tiles, double* part, double* lp_out,
                             double* mu_out) {
  PGB_CUDA(cudaFuncSetAttribute(stats_pass_kernel<HAS_Y, FIN, FAMK, RPT, OTHER, REP, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kSmemMax));
  stats_pass_kernel<HAS_Y, FIN, FAMK, RPT, OTHER, REP, CT><<<grid, CT + 32, smem, st>>>(
      P, M->d_terms, M->n_terms, M->m, M->n_features, M->fam, X, n, ld, y, w, beta, scale, ybar, poisson_round, NS,
      n_bulk_tiles, part, lp_out, mu_out);
  return PGB_OK;
}

// everything one launch needs, filled by pass_run (pgb_pass.cu)
struct PassLaunch {
  PassPlan P;
  const pgb_model* M;
  int grid, NS, RPT, famk, threads;
  size_t smem;
  cudaStream_t st;
  const double* X;
  int64_t n, ld, n_bulk_tiles;
  const double* y;
  const float* w;
  const double* beta;
  double scale, ybar;
  int poisson_round;
  double *part, *lp_out, *mu_out;
};

// the template instance for (HAS_Y, FIN) x family x rows per thread x general terms x table copies
template <bool HAS_Y, int FIN>
int pass_dispatch(const PassLaunch& A) {
  const PassPlan& P = A.P;
  int rc = PGB_OK;
#define PGB_STATS_ARGS P, A.M, A.grid, A.smem, A.st, A.X, A.n, A.ld, A.y, A.w, A.beta, A.scale, A.ybar, A.poisson_round, A.NS, A.n_bulk_tiles, A.part, A.lp_out, A.mu_out
#define PGB_STATS_REP(FK, RP, OT)                                                                                   \
  if constexpr (!(OT)) {                                                                                            \
    if (A.threads == kPassThreadsWide) {  /* pass_run: only lean models with eight table copies */                  \
      rc = stats_pass_launch<HAS_Y, FIN, FK, RP, OT, 8, kPassThreadsWide>(PGB_STATS_ARGS);                          \
      break;                                                                                                        \
    }                                                                                                               \
  }                                                                                                                 \
  rc = P.rep_shift ? stats_pass_launch<HAS_Y, FIN, FK, RP, OT, 8, kPassThreads>(PGB_STATS_ARGS)                     \
                   : stats_pass_launch<HAS_Y, FIN, FK, RP, OT, 1, kPassThreads>(PGB_STATS_ARGS);
#define PGB_STATS_FAM(RP, OT)                        \
  switch (A.famk) {                                  \
    case 1: PGB_STATS_REP(1, RP, OT) break;          \
    case 2: PGB_STATS_REP(2, RP, OT) break;          \
    case 3: PGB_STATS_REP(3, RP, OT) break;          \
    default: PGB_STATS_REP(0, RP, OT) break;         \
  }
  bool done = false;
  if (P.n_other) {
    PGB_STATS_FAM(1, true)
    done = true;
  }
  if constexpr (FIN != PASS_FULL) {  // two rows per thread: never with the log-likelihood sums (registers)
    if (!done && A.RPT == 2) {
      PGB_STATS_FAM(2, false)
      done = true;
    }
  }
  if (!done) { PGB_STATS_FAM(1, false) }
#undef PGB_STATS_FAM
#undef PGB_STATS_REP
#undef PGB_STATS_ARGS
  return rc;
}
