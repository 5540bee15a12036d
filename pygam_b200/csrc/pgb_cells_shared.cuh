// pgb_cells_shared.cuh -- row-level arithmetic of the cell-owned Gram pass that is shared between the
// kernels of pgb_cells.cu and the host-side self check of the plan (pgb_selfcheck.cu, tests only): the cell
// record of a (row, marginal), the moment "powers" of a record, one row's contribution to the sums of a
// generic pair, the cell key of a row, the change of basis moments -> B-spline entries along one axis and
// the gather of one entry of [G | r] from the sums.
#pragma once
#include "pgb_internal.h"

// the four cubic B-splines of an interval as polynomials of the centred local coordinate s = t - 1/2:
// b_a(s) = sum_p kMh[a][p] s^p
#define PGB_M_ROWS                                                                                    \
  {{1.0 / 48.0, -1.0 / 8.0, 1.0 / 4.0, -1.0 / 6.0}, {23.0 / 48.0, -5.0 / 8.0, -1.0 / 4.0, 1.0 / 2.0}, \
   {23.0 / 48.0, 5.0 / 8.0, -1.0 / 4.0, -1.0 / 2.0}, {1.0 / 48.0, 1.0 / 8.0, 1.0 / 4.0, 1.0 / 6.0}}

// eval_marg's branch outside [0, 1] for a cubic marginal with `nint` intervals (utils.py:744-763), from the
// normalised coordinate u
PGB_HD void extrap_basis_n(int nint, double u, double* v) {
  const bool below = u < 0.0;
  const double te = below ? 0.0 : 1.0, dist = below ? u : u - 1.0;
  double lowv[PGB_MAXV];
  deboor_uniform<2>(te, lowv);
  deboor_uniform<3>(te, v);
#pragma unroll
  for (int r = 0; r <= 3; ++r) {
    const double a = (r >= 1) ? lowv[r - 1] : 0.0;
    const double b = (r <= 2) ? lowv[r] : 0.0;
    v[r] = ((double)nint * (a - b)) * dist + v[r];
  }
}

// cell record of one cubic marginal: knot interval (bit 7: outside the edge knots), the record coordinate
// (centred local coordinate in [-1/2, 1/2]; outside the knots the normalised coordinate moved out of that
// range) and the four basis values
PGB_HD void marg_record(const DMarg& mg, double x, int& cell, double& t, double* v) {
  const double u = (x - mg.lo) * mg.inv_span;
  if (u >= 0.0 && u <= 1.0) {  // the arithmetic of eval_term's straight-line cubic
    const double s = u * (double)mg.nint;
    cell = min((int)s, mg.nint - 1);
    t = s - (double)cell;
    cubic_bspline(t, v[0], v[1], v[2], v[3]);
    t -= 0.5;
  } else {
    extrap_basis_n(mg.nint, u, v);
    cell = (u < 0.0 ? 0 : mg.nint - 1) | 0x80;
    t = (u < 0.0) ? u - 1.0 : u + 1.0;
  }
}

// lp of one row over every term kind of a cell-eligible model, storing the row's cell records on the way:
// cubic marginals (s(), both marginals of te()): interval byte + record coordinate; f(): the coefficient index of
// the row's level (0xff: none); l(): the feature value.  te() rows follow TensorTerm.build_columns
// (terms.py:1454-1477, utils.py:899-948: index a * K_2 + b), f() rows FactorTerm.build_columns
// (terms.py:1077-1096), l() rows LinearTerm.build_columns (terms.py:693-708).
template <class LoadX, class StoreT, class StoreC>
PGB_HD double prepass_row_lp(const DTerm* terms, int n_terms, const double* beta, const LoadX& x, const StoreT& store_t,
                             const StoreC& store_c) {
  double lp = 0.0;
  for (int q = 0; q < n_terms; ++q) {
    const DTerm& T = terms[q];
    const double* bt = beta + T.coef_offset;
    if (T.kind == PGB_TERM_INTERCEPT) {
      lp += bt[0];
    } else if (T.kind == PGB_TERM_LINEAR) {
      const double xv = x(T.marg[0].feature);
      lp = fma(xv, bt[0], lp);
      store_t(T.fc0, xv);
    } else if (T.kind == PGB_TERM_FACTOR) {
      double v[PGB_MAXV];
      const int i0 = eval_marg(T.marg[0], x(T.marg[0].feature), v);
      const int c = i0 - T.drop_first;
      const bool ok = v[0] != 0.0 && c >= 0;
      if (ok) lp += bt[c];
      store_c(T.fc0, ok ? c : 0xff);
    } else if (T.kind == PGB_TERM_SPLINE) {
      double v[4], t;
      int cell;
      marg_record(T.marg[0], x(T.marg[0].feature), cell, t, v);
      bt += cell & 0x7f;
      if (T.by >= 0) {  // s(j, by=k): the basis row times x_k (terms.py:953-954); x_k is the record of a value column
        const double byv = x(T.by);
        lp = fma(v[0] * byv, bt[0], lp);
        lp = fma(v[1] * byv, bt[1], lp);
        lp = fma(v[2] * byv, bt[2], lp);
        lp = fma(v[3] * byv, bt[3], lp);
        store_t(T.fc0 + 1, byv);
      } else {
        lp = fma(v[0], bt[0], lp);
        lp = fma(v[1], bt[1], lp);
        lp = fma(v[2], bt[2], lp);
        lp = fma(v[3], bt[3], lp);
      }
      store_t(T.fc0, t);
      store_c(T.fc0, cell);
    } else if (T.marg[0].order < 0 || T.marg[1].order < 0) {  // te(s, l) / te(l, s): one cubic and one l() marginal
      const int qs = T.marg[0].order < 0 ? 1 : 0, ql = 1 - qs;
      double v[4], t;
      int cell;
      marg_record(T.marg[qs], x(T.marg[qs].feature), cell, t, v);
      const double xl = x(T.marg[ql].feature);
      bt += cell & 0x7f;  // the l() marginal has a single column: index a * 1 + 0 or 0 * K + a
#pragma unroll
      for (int a = 0; a < 4; ++a) lp = fma(v[a] * xl, bt[a], lp);
      store_t(T.fc0 + qs, t);
      store_c(T.fc0 + qs, cell);
      store_t(T.fc0 + ql, xl);
    } else {  // te() of two cubic marginals
      double va[4], vb[4], t0, t1;
      int c0, c1;
      marg_record(T.marg[0], x(T.marg[0].feature), c0, t0, va);
      marg_record(T.marg[1], x(T.marg[1].feature), c1, t1, vb);
      const int K1 = T.marg[1].nspl;
      bt += (c0 & 0x7f) * K1 + (c1 & 0x7f);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) lp = fma(va[a] * vb[b], bt[a * K1 + b], lp);
      store_t(T.fc0, t0);
      store_c(T.fc0, c0);
      store_t(T.fc0 + 1, t1);
      store_c(T.fc0 + 1, c1);
    }
  }
  return lp;
}

// "powers" of a record coordinate: (1, s, s^2, s^3) inside the knots; for a row outside the knots (rare: custom
// edge_knots narrower than the data) mu = M^-1 v of its four extrapolated basis values, so that M mu = v and
// the moment sums stay valid
PGB_HD void record_powers(int nint, double s, double* p) {
  if (fabs(s) > 0.75) {
    const double u = s < 0.0 ? s + 1.0 : s - 1.0;
    double v[PGB_MAXV];
    extrap_basis_n(nint, u, v);
    p[0] = v[0] + v[1] + v[2] + v[3];
    p[1] = 1.5 * (v[3] - v[0]) + 0.5 * (v[2] - v[1]);
    p[2] = (23.0 / 12.0) * (v[0] + v[3]) - (1.0 / 12.0) * (v[1] + v[2]);
    p[3] = 1.875 * (v[3] - v[0]) + 0.375 * (v[1] - v[2]);
  } else {
    p[0] = 1.0;
    p[1] = s;
    p[2] = s * s;
    p[3] = p[2] * s;
  }
}

PGB_HD double sel4(const double* p, int e) { return e == 0 ? p[0] : e == 1 ? p[1] : e == 2 ? p[2] : p[3]; }

// one row of a generic pair for the thread of slice `sg`: acc[4 p + q] += (weight * outer powers * values *
// a_p) * b_q.  `col(c)` returns the row's record coordinate of factor column c, `wgt` is omega (or omega z).
template <class Col>
PGB_HD void gen_row_acc(const DGenPair& P, int sg, double wgt, const Col& col, double* acc) {
  double mult = wgt;
  for (int q = 0; q < P.nval; ++q) mult *= col(P.val_col[q]);
  if (P.out0 >= 0) {
    double pw[4];
    record_powers(P.f_n[P.out0], col(P.f_col[P.out0]), pw);
    mult *= sel4(pw, sg & 3);
    if (P.out1 >= 0) {
      record_powers(P.f_n[P.out1], col(P.f_col[P.out1]), pw);
      mult *= sel4(pw, (sg >> 2) & 3);
    }
  }
  if (P.in_b < 0) {
    acc[0] += mult;
    return;
  }
  double b[4];
  record_powers(P.f_n[P.in_b], col(P.f_col[P.in_b]), b);
  if (P.in_a < 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = fma(mult, b[q], acc[q]);
    return;
  }
  double a[4];
  record_powers(P.f_n[P.in_a], col(P.f_col[P.in_a]), a);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const double u = mult * a[p];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[4 * p + q] = fma(u, b[q], acc[4 * p + q]);
  }
}

// cell key of tile row r of an order, 0xffff: the row has no entry (a level byte 0xff: the dropped level of
// f(coding="dummy"), or a value outside the compiled levels).  `cellbyte(c)` returns the row's byte of column c.
template <class Cell>
PGB_HD uint32_t gen_key(const DGenOrder& O, int r, const Cell& cellbyte) {
  uint32_t key = 0;
  for (int d = 0; d < O.nkey; ++d) {
    const uint32_t c = cellbyte(O.key_col[d]);
    if (O.key_kind[d] == PGB_FK_LEVEL) {
      if (c >= (uint32_t)O.key_n[d]) return 0xffffu;
      key += c * (uint32_t)O.key_mul[d];
    } else {
      key += (c & 0x7fu) * (uint32_t)O.key_mul[d];
    }
  }
  return key * (uint32_t)O.nh + (((uint32_t)r * O.hmul) >> 16);
}

// every (product, cell) of a generic pair that contributes to the entry (local coefficient la of term ta, lb of
// term tb): f(offset of the first of the nh hash splits inside `sums`)
template <class F>
PGB_HD void gen_visit(const DGenPair& P, int la, int lb, F&& f) {
  int idx[4];
  for (int q = 0; q < P.nf; ++q) {
    const int l = P.f_side[q] ? lb : la;
    idx[q] = (l / P.f_stride[q]) % P.f_nspl[q];
  }
  const int ncomb = 1 << (2 * P.n_ax);
  for (int c = 0; c < ncomb; ++c) {
    int cellk[4] = {-1, -1, -1, -1};
    bool ok = true;
    for (int q = 0; q < P.nf && ok; ++q) {
      int cf;
      if (P.f_kind[q] == PGB_FK_CUBIC) {
        cf = idx[q] - ((c >> (2 * P.f_ax[q])) & 3);
        if (cf < 0 || cf >= P.f_n[q]) ok = false;
      } else if (P.f_kind[q] == PGB_FK_LEVEL) {
        cf = idx[q];
      } else {
        continue;
      }
      const int kd = P.f_key[q];
      if (cellk[kd] < 0) cellk[kd] = cf;
      else if (cellk[kd] != cf) ok = false;
    }
    if (!ok) continue;
    int64_t key = 0;
    for (int d = 0; d < 4; ++d)
      if (cellk[d] >= 0) key += (int64_t)cellk[d] * P.key_mul[d];
    f(P.sum_off + (int64_t)c * P.ncells + key * P.nh);
  }
}

// change of basis along one moment axis: the four sums x_d -> sum_d M[a][d] x_d
PGB_HD void moment_line(double* x0, int64_t stride) {
  const double kMl[4][4] = PGB_M_ROWS;
  const double v0 = x0[0], v1 = x0[stride], v2 = x0[2 * stride], v3 = x0[3 * stride];
#pragma unroll
  for (int a = 0; a < 4; ++a) x0[a * stride] = kMl[a][0] * v0 + kMl[a][1] * v1 + kMl[a][2] * v2 + kMl[a][3] * v3;
}

// moment sums of one cell of a spline pair -> the entries of G the cell touches, in place: B = M S M^T (pair), the
// packed upper triangle of M S M^T with S symmetric (diagonal block; sums 10..13: M x for the intercept column,
// 14..17: for the right-hand side).  S[e * st] is sum e of the cell.
PGB_HD void spline_transform_cell(bool diag, double* S, int64_t st) {
  const double kMl[4][4] = PGB_M_ROWS;
  double x[4][4], y[4][4];
  if (!diag) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) x[p][q] = S[(4 * p + q) * st];
  } else {
    int e = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = p; q < 4; ++q) { x[p][q] = S[e * st]; x[q][p] = x[p][q]; ++e; }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int q = 0; q < 4; ++q) y[a][q] = kMl[a][0] * x[0][q] + kMl[a][1] * x[1][q] + kMl[a][2] * x[2][q] + kMl[a][3] * x[3][q];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) x[a][b] = y[a][0] * kMl[b][0] + y[a][1] * kMl[b][1] + y[a][2] * kMl[b][2] + y[a][3] * kMl[b][3];
  if (!diag) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) S[(4 * a + b) * st] = x[a][b];
  } else {
    int e = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) S[(e++) * st] = x[a][b];
#pragma unroll
    for (int g = 10; g <= 14; g += 4) {
      const double v0 = S[g * st], v1 = S[(g + 1) * st], v2 = S[(g + 2) * st], v3 = S[(g + 3) * st];
#pragma unroll
      for (int a = 0; a < 4; ++a) S[(g + a) * st] = kMl[a][0] * v0 + kMl[a][1] * v1 + kMl[a][2] * v2 + kMl[a][3] * v3;
    }
  }
}

// Which sums make up entry (r, c), r <= c <= m (c == m: the right-hand side), of [G | r]: f(offset, count) for every
// run sums[offset .. offset + count) that is added (count > 1: the hash splits of a cell); offset -1 / -2: the
// prepass sums of omega / omega z (the intercept's own entries).  Shared by the assembly kernel and the host-side
// coverage check.
struct CellsView {
  const int32_t *coef_term, *pair_of, *term_class, *gpair_of;
  const DCellPair* pairs;
  const DGenPair* gpairs;
  int nt, m, icpt;
};
template <class F>
PGB_HD void cells_entry_visit(const CellsView& V, int r, int c, F&& f) {
  const int m = V.m, nt = V.nt, icpt = V.icpt;
  const int tr = V.coef_term[r], tc = (c < m) ? V.coef_term[c] : nt;
  const int clr = V.term_class[tr], clc = (c < m) ? V.term_class[tc] : 0;
  if (clr == 2 || clc == 2) {
    const int ta = tr < tc ? tr : tc, tb = tr < tc ? tc : tr;
    const DGenPair& P = V.gpairs[V.gpair_of[(size_t)ta * (nt + 1) + tb]];
    int la, lb;
    if (tr <= tc) { la = r - P.off_a; lb = (c < m) ? c - P.off_b : 0; }
    else { la = c - P.off_a; lb = r - P.off_b; }
    const int nh = P.nh;
    gen_visit(P, la, lb, [&](int64_t off) { f(off, nh); });
    return;
  }
  if (r == icpt && (c == m || c == icpt)) {
    f((int64_t)(c == m ? -2 : -1), 0);
  } else if (c == m || c == icpt || r == icpt) {
    const int cs = (r == icpt) ? c : r;  // the spline coefficient
    const int ts = V.coef_term[cs];
    const DCellPair& P = V.pairs[V.pair_of[(size_t)ts * nt + ts]];
    const int p = cs - P.off_i;
    for (int a = 0; a < 4; ++a) {
      const int ci = p - a;
      if (ci < 0 || ci >= P.nint_i) continue;
      f(P.sum_off + (int64_t)((c == m ? 14 : 10) + a) * P.ncells + ci * P.ncb, P.ncb);
    }
  } else if (tr == tc) {
    const DCellPair& P = V.pairs[V.pair_of[(size_t)tr * nt + tr]];
    const int p = r - P.off_i, d = c - r;
    if (d > 3) return;
    for (int a = 0; a + d < 4; ++a) {
      const int ci = p - a;
      if (ci < 0 || ci >= P.nint_i) continue;
      const int e = a * 4 - a * (a - 1) / 2 + d;  // (a, a + d) of the packed upper triangle
      f(P.sum_off + (int64_t)e * P.ncells + ci * P.ncb, P.ncb);
    }
  } else {
    const int ta = tr < tc ? tr : tc, tb = tr < tc ? tc : tr;
    const DCellPair& P = V.pairs[V.pair_of[(size_t)ta * nt + tb]];
    const int p = ((tr == ta) ? r : c) - P.off_i, q = ((tr == ta) ? c : r) - P.off_j;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        const int ci = p - a, cj = q - b;
        if (ci < 0 || ci >= P.nint_i || cj < 0 || cj >= P.ncb) continue;
        f(P.sum_off + (int64_t)(4 * a + b) * P.ncells + ci * P.ncb + cj, 1);
      }
  }
}

