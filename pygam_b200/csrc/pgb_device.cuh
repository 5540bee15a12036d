// pgb_device.cuh -- device-side model descriptor, on-the-fly basis evaluation and the
// link / distribution step shared by every kernel of the engine.
//
// What it replaces in the reference (pyGAM 0.12.0): utils.b_spline_basis (utils.py:624-770),
// utils.tensor_product (utils.py:899-948), Term.build_columns (terms.py:584-599, 693-708,
// 926-956, 1077-1096, 1454-1477), Link.mu/gradient/link (links.py) and Distribution.V /
// deviance / log_pdf (distributions.py).  Nothing here materialises the model matrix: a row
// of it is produced as (column, value) pairs handed to an `emit` functor.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pgb.h"

#define PGB_SQRT_EPS 1.4901161193847656e-08 /* sqrt(np.finfo(float).eps), pygam.py:657, 749 */
#define PGB_MAXV (PGB_MAX_ORDER + 1)
// host + device: the row-level arithmetic is shared with the host-side self check of the cell plan (tests only)
#define PGB_HD __host__ __device__ __forceinline__

struct DMarg {
  int32_t feature, nspl, order, periodic;
  int32_t nint;  // uniform intervals on the normalised range: nspl(+order if periodic) - order
  int32_t pad;
  double lo, span, inv_span;
  double knot_step;  // 1/nint as numpy.linspace computes it (utils.py:686)
};

struct DTerm {
  int32_t kind, n_marg, by, drop_first, coef_offset, n_coefs, nnz, slot0;
  // placement of the term inside the Gram plan (pgb_internal.h): coefficient group, index of the
  // term's first coefficient inside the group, first global record slot (4 per quad)
  int32_t group, idx_base, gslot0;
  int32_t fast;  // plain cubic spline term (order 3, basis "ps", no by): straight-line evaluation
  int32_t fc0;   // first column of the term's factors in the cell records (pgb_cells.cu), -1: none
  int32_t pad0;
  DMarg marg[PGB_MAX_MARG];
};

struct Family {
  int32_t dist, link;
  double levels, expectile;
};

// ----------------------------------------------------------------------------------------
// uniform B-spline values on one interval: Cox-de Boor with unit knot spacing (the reference
// runs the same recursion on its explicit knot vector, utils.py:717-734).  v[r], r = 0..D, is
// the value of basis function (interval + r) at local coordinate t in [0, 1].
// ----------------------------------------------------------------------------------------
template <int D>
PGB_HD void deboor_uniform(double t, double* v) {
  v[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= D; ++j) {
    const double inv = 1.0 / (double)j;
    double saved = 0.0;
#pragma unroll
    for (int r = 0; r < j; ++r) {
      // left = t + (j - r - 1), right = (r + 1) - t, left + right = j
      const double tmp = v[r] * inv;
      v[r] = saved + ((double)(r + 1) - t) * tmp;
      saved = (t + (double)(j - r - 1)) * tmp;
    }
    v[j] = saved;
  }
}

// the four non-zero cubic B-splines of a uniform knot vector at local coordinate t, closed form:
// [(1-t)^3, 3t^3 - 6t^2 + 4, -3t^3 + 3t^2 + 3t + 1, t^3] / 6; agrees with
// the recursion above to rounding and is what the straight-line cubic path uses everywhere
PGB_HD void cubic_bspline(double t, double& b0, double& b1, double& b2, double& b3) {
  const double omt = 1.0 - t, t2 = t * t, t3 = t2 * t;
  b0 = (omt * omt) * (omt * (1.0 / 6.0));
  b3 = t3 * (1.0 / 6.0);
  b1 = fma(t3, 0.5, fma(t2, -1.0, 2.0 / 3.0));
  b2 = fma(t3, -0.5, fma(t2, 0.5, fma(t, 0.5, 1.0 / 6.0)));
}

PGB_HD void deboor_dispatch(int d, double t, double* v) {
  switch (d) {
    case 0: v[0] = 1.0; break;
    case 1: deboor_uniform<1>(t, v); break;
    case 2: deboor_uniform<2>(t, v); break;
    case 3: deboor_uniform<3>(t, v); break;
    case 4: deboor_uniform<4>(t, v); break;
    case 5: deboor_uniform<5>(t, v); break;
    case 6: deboor_uniform<6>(t, v); break;
    default: deboor_uniform<7>(t, v); break;
  }
}

// One marginal basis row.  Returns the index of the first non-zero basis function; v[0..order]
// are the values (always order+1 entries, zeros included, so that every row of a term has the
// same number of slots).  Column indices of a periodic basis wrap modulo nspl in the caller.
PGB_HD int eval_marg(const DMarg& mg, double x, double* v) {
  const int d = mg.order;
  if (d == 0) {
    // order-0 (factor / Haar) basis is discontinuous: pick the cell with the reference's own
    // comparisons against linspace knots, last knot nudged by 1e-9 (utils.py:686-714)
    double u = (x - mg.lo) / mg.span;
    if (mg.periodic) {
      u = fmod(u, 1.0 + 1e-9);
      if (u < 0.0) u += 1.0 + 1e-9;
    }
    const int K = mg.nint;
    v[0] = 0.0;
    if (!(u >= 0.0) || !(u < 1.0 + 1e-9)) return 0;
    int i = (int)(u * (double)K);
    i = min(max(i, 0), K - 1);
    if (i > 0 && u < (double)i * mg.knot_step) --i;
    else if (i < K - 1 && u >= (double)(i + 1) * mg.knot_step) ++i;
    v[0] = 1.0;
    return i;
  }
  double u = (x - mg.lo) * mg.inv_span;
  if (mg.periodic) {  // utils.py:693-694
    u = fmod(u, 1.0 + 1e-9);
    if (u < 0.0) u += 1.0 + 1e-9;
  }
  const int nint = mg.nint;
  if (u < 0.0 || u > 1.0) {
    // linear extrapolation from the edge (utils.py:744-763): value at the edge plus slope
    // times distance; the slope of a degree-d B-spline is nint * (N_{j,d-1} - N_{j+1,d-1})
    const bool below = u < 0.0;
    const double te = below ? 0.0 : 1.0;
    const double dist = below ? u : u - 1.0;
    double lowv[PGB_MAXV];
    deboor_dispatch(d - 1, te, lowv);
    deboor_dispatch(d, te, v);
#pragma unroll
    for (int r = 0; r < PGB_MAXV; ++r) {
      if (r <= d) {
        const double a = (r >= 1) ? lowv[r - 1] : 0.0;
        const double b = (r <= d - 1) ? lowv[r] : 0.0;
        v[r] = ((double)nint * (a - b)) * dist + v[r];
      }
    }
    return below ? 0 : nint - 1;
  }
  const double s = u * (double)nint;
  int i = (int)s;
  i = min(i, nint - 1);
  const double t = s - (double)i;
  deboor_dispatch(d, t, v);
  return i;
}

// ----------------------------------------------------------------------------------------
// one model-matrix row of one term as (column, value) pairs; exactly T.nnz calls of emit
// ----------------------------------------------------------------------------------------
// `load(f)` returns feature f of the row being evaluated: straight from HBM (LoadGlobal) or from a
// tile staged in shared memory by bulk async copies (LoadTile)
struct LoadGlobal {
  const double* __restrict__ X;
  int64_t ld, row;
  __device__ __forceinline__ double operator()(int f) const { return __ldg(X + (int64_t)f * ld + row); }
  __device__ __forceinline__ double at(int f, int) const { return __ldg(X + (int64_t)f * ld + row); }
};
struct LoadTile {
  const double* tile;  // [n_features][tile_ld], element (f, r) at tile[f * tile_ld + r]
  int tile_ld, r;
  __device__ __forceinline__ double operator()(int f) const { return tile[f * tile_ld + r]; }
  // the same element through its precomputed byte offset f * tile_ld * 8
  __device__ __forceinline__ double at(int, int byte_off) const {
    return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tile + r) + byte_off);
  }
};

template <class Emit, class Load>
__device__ __forceinline__ void eval_term(const DTerm& T, const Load& load, Emit& emit) {
  if (T.fast) {
    // the common case, s(feature) with the default cubic P-spline basis, inside the edge knots:
    // same arithmetic as the general path below, without its dispatch
    const DMarg& mg = T.marg[0];
    const double u = (load(mg.feature) - mg.lo) * mg.inv_span;
    if (u >= 0.0 && u <= 1.0) {
      const double s = u * (double)mg.nint;
      const int i = min((int)s, mg.nint - 1);
      double b0, b1, b2, b3;
      cubic_bspline(s - (double)i, b0, b1, b2, b3);
      const int c = T.coef_offset + i;
      emit(c, b0);
      emit(c + 1, b1);
      emit(c + 2, b2);
      emit(c + 3, b3);
      return;
    }
  }
  if (T.kind == PGB_TERM_INTERCEPT) {
    emit(T.coef_offset, 1.0);
    return;
  }
  if (T.kind == PGB_TERM_LINEAR) {
    emit(T.coef_offset, load(T.marg[0].feature));
    return;
  }
  const double byv = (T.by >= 0) ? load(T.by) : 1.0;
  if (T.n_marg == 1) {
    const DMarg& mg = T.marg[0];
    double v[PGB_MAXV];
    const int i0 = eval_marg(mg, load(mg.feature), v);
    const int K = mg.nspl;
#pragma unroll
    for (int a = 0; a < PGB_MAXV; ++a) {
      if (a <= mg.order) {
        int c = i0 + a;
        if (c >= K) c -= K;  // cyclic fold (utils.py:736-742); never taken for basis "ps"
        double val = v[a] * byv;
        if (T.drop_first) {  // f(coding="dummy"), terms.py:1093-1094
          c -= 1;
          if (c < 0) { c = 0; val = 0.0; }
        }
        emit(T.coef_offset + c, val);
      }
    }
    return;
  }
  // tensor term: row-wise Kronecker product, first marginal slowest (utils.py:943-946)
  double v[PGB_MAX_MARG][PGB_MAXV];
  int i0[PGB_MAX_MARG], cnt[PGB_MAX_MARG], K[PGB_MAX_MARG];
#pragma unroll
  for (int q = 0; q < PGB_MAX_MARG; ++q) {
    if (q < T.n_marg) {
      const DMarg& mg = T.marg[q];
      if (mg.order < 0) {  // l() marginal: one column holding the feature value (terms.py:693-708)
        v[q][0] = load(mg.feature);
        i0[q] = 0;
        cnt[q] = 1;
        K[q] = 1;
      } else {
        i0[q] = eval_marg(mg, load(mg.feature), v[q]);
        cnt[q] = mg.order + 1;
        K[q] = mg.nspl;
      }
    } else {
      i0[q] = 0; cnt[q] = 1; K[q] = 1; v[q][0] = 1.0;
    }
  }
  for (int a0 = 0; a0 < cnt[0]; ++a0) {
    int c0 = i0[0] + a0; if (c0 >= K[0]) c0 -= K[0];
    const double p0 = v[0][a0] * byv;
    for (int a1 = 0; a1 < cnt[1]; ++a1) {
      int c1 = i0[1] + a1; if (c1 >= K[1]) c1 -= K[1];
      const double p1 = p0 * v[1][a1];
      const int c01 = c0 * K[1] + c1;
      for (int a2 = 0; a2 < cnt[2]; ++a2) {
        int c2 = i0[2] + a2; if (c2 >= K[2]) c2 -= K[2];
        const double p2 = p1 * v[2][a2];
        const int c012 = c01 * K[2] + c2;
        for (int a3 = 0; a3 < cnt[3]; ++a3) {
          int c3 = i0[3] + a3; if (c3 >= K[3]) c3 -= K[3];
          emit(T.coef_offset + c012 * K[3] + c3, p2 * v[3][a3]);
        }
      }
    }
  }
}

template <class Emit>
__device__ __forceinline__ void eval_term(const DTerm& T, const double* __restrict__ X, int64_t ld,
                                          int64_t row, Emit& emit) {
  const LoadGlobal load{X, ld, row};
  eval_term(T, load, emit);
}

// ----------------------------------------------------------------------------------------
// links (links.py:21-314)
// ----------------------------------------------------------------------------------------
PGB_HD double link_inv(const Family& f, double lp) {
  switch (f.link) {
    case PGB_LINK_IDENTITY: return lp;
    case PGB_LINK_LOGIT: { const double e = exp(lp); return f.levels * e / (e + 1.0); }  // links.py:121-122
    case PGB_LINK_LOG: return exp(lp);
    case PGB_LINK_INVERSE: return 1.0 / lp;
    default: return 1.0 / sqrt(lp);
  }
}
PGB_HD double link_fwd(const Family& f, double mu) {
  switch (f.link) {
    case PGB_LINK_IDENTITY: return mu;
    case PGB_LINK_LOGIT: return log(mu) - log(f.levels - mu);  // links.py:105
    case PGB_LINK_LOG: return log(mu);
    case PGB_LINK_INVERSE: return 1.0 / mu;
    default: return 1.0 / (mu * mu);
  }
}
PGB_HD double link_grad(const Family& f, double mu) {
  switch (f.link) {
    case PGB_LINK_IDENTITY: return 1.0;
    case PGB_LINK_LOGIT: return f.levels / (mu * (f.levels - mu));  // links.py:137
    case PGB_LINK_LOG: return 1.0 / mu;
    case PGB_LINK_INVERSE: return -1.0 / (mu * mu);
    default: return -2.0 / (mu * mu * mu);
  }
}

// ----------------------------------------------------------------------------------------
// distributions (distributions.py:97-612)
// ----------------------------------------------------------------------------------------
PGB_HD double dist_var(const Family& f, double mu) {
  switch (f.dist) {
    case PGB_DIST_NORMAL: return 1.0;
    case PGB_DIST_BINOMIAL: return mu * (1.0 - mu / f.levels);
    case PGB_DIST_POISSON: return mu;
    case PGB_DIST_GAMMA: return mu * mu;
    default: return mu * mu * mu;
  }
}
PGB_HD double ylogydu(double y, double u) {  // utils.py:773-789
  return (y != 0.0) ? y * log(y / u) : 0.0;
}
// unscaled, unweighted deviance of one row
PGB_HD double dist_dev(const Family& f, double y, double mu) {
  switch (f.dist) {
    case PGB_DIST_NORMAL: { const double r = y - mu; return r * r; }
    case PGB_DIST_BINOMIAL:
      // 0/1 responses: 2 log(1/mu) or 2 log(1/(1-mu)) is all that is left of the general formula
      if (f.levels == 1.0 && (y == 0.0 || y == 1.0)) return -2.0 * log((y == 1.0) ? mu : 1.0 - mu);
      return 2.0 * (ylogydu(y, mu) + ylogydu(f.levels - y, f.levels - mu));
    case PGB_DIST_POISSON: return 2.0 * (ylogydu(y, mu) - (y - mu));
    case PGB_DIST_GAMMA: return 2.0 * ((y - mu) / mu - log(y / mu));
    default: { const double r = y - mu; return (r * r) / (mu * mu * y); }
  }
}

#define PGB_HALF_LOG_2PI 0.91893853320467274178
// log pdf / pmf of one row as scipy.stats evaluates it for the reference
// (distributions.py:110-131, 225-247, 323-352, 428-449, 531-552)
__device__ __forceinline__ double dist_logpdf(const Family& f, double y, double mu, double w,
                                              double scale) {
  switch (f.dist) {
    case PGB_DIST_NORMAL: {
      const double sd = scale / w;
      const double zz = (y - mu) / sd;
      return -0.5 * zz * zz - PGB_HALF_LOG_2PI - log(sd);
    }
    case PGB_DIST_BINOMIAL: {
      const double n = f.levels, p = mu / f.levels;
      const double comb = lgamma(n + 1.0) - lgamma(y + 1.0) - lgamma(n - y + 1.0);
      const double a = (y != 0.0) ? y * log(p) : 0.0;             // xlogy
      const double b = (n - y != 0.0) ? (n - y) * log1p(-p) : 0.0;  // xlog1py
      return comb + a + b;
    }
    case PGB_DIST_POISSON: {
      const double lam = mu * w;
      const double a = (y != 0.0) ? y * log(lam) : 0.0;
      return a - lam - lgamma(y + 1.0);
    }
    case PGB_DIST_GAMMA: {
      const double nu = w / scale;
      const double th = mu / nu;
      const double xs = y / th;
      const double a = (nu - 1.0 != 0.0) ? (nu - 1.0) * log(xs) : 0.0;
      return a - xs - lgamma(nu) - log(th);
    }
    default: {  // invgauss.logpdf(y, mu, scale=s), s = scale / w
      const double s = scale / w;
      const double xs = y / s;
      const double r = xs - mu;
      return -PGB_HALF_LOG_2PI - 1.5 * log(xs) - (r * r) / (2.0 * xs * mu * mu) - log(s);
    }
  }
}

// ----------------------------------------------------------------------------------------
// the elementwise IRLS step of one row (pygam.py:764-777, 580-663, 3419-3460)
// ----------------------------------------------------------------------------------------
struct IrlsRow {
  double sw;     // W (the square root of the IRLS weight) of a kept row, 0 for a masked row
  double omega;  // W^2 of the kept row, 0 for a masked row
  double z;      // working response lp + (y - mu) g'(mu), 0 for a masked row
  double mu;
  bool keep;
};

// winv is 1 / w as the reference computes it: user weights are float32 (pygam.py:889), so
// `weights ** -1` (pygam.py:633) is a float32 reciprocal; 1.0 when no weights were given
PGB_HD IrlsRow irls_row(const Family& f, double lp, double y, double winv) {
  IrlsRow r;
  const double mu = link_inv(f, lp);
  const double g = link_grad(f, mu);
  const double V = dist_var(f, mu);
  double Wd = 1.0 / sqrt(g * g * V * winv);  // (g'^2 V / w)^-0.5, pygam.py:631-636
  if (f.expectile > 0.0) {                         // ExpectileGAM._W, pygam.py:3449-3460
    const double asym = (y > mu) ? f.expectile : 1.0 - f.expectile;
    Wd *= sqrt(asym);
  }
  r.keep = (fabs(Wd) >= PGB_SQRT_EPS) && isfinite(Wd);  // GAM._mask, pygam.py:657
  r.mu = mu;
  r.sw = r.keep ? Wd : 0.0;
  r.omega = r.keep ? Wd * Wd : 0.0;
  r.z = r.keep ? lp + (y - mu) * g : 0.0;  // GAM._pseudo_data, pygam.py:597
  return r;
}

PGB_HD double weight_inv(const float* __restrict__ w, int64_t i) {
#ifdef __CUDA_ARCH__
  return w ? (double)__fdiv_rn(1.0f, __ldg(w + i)) : 1.0;
#else
  return w ? (double)(1.0f / w[i]) : 1.0;
#endif
}

// response of _initial_estimate (pygam.py:696-704)
PGB_HD double init_response(const Family& f, double y) {
  if (y == 0.0) y += 0.01;
  if (y == 1.0) y -= 0.01;
  return link_fwd(f, y);
}
