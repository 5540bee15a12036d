// pgb_selfcheck.cu -- TEST INFRASTRUCTURE, not part of libpgb.so: a host-side walk through the cell decomposition of
// the Gram pass (pgb_cells.cu) with the SAME plan (pgb_plan_cells), row arithmetic (pgb_cells_shared.cuh: records,
// keys, moment sums, change of basis, assembly) and sums layout as the kernels, one row at a time.  Built into
// libpgb_selfcheck.so by build.sh / __graft_entry__.build(); tests/test_cells_host.py compares its [G | r] with the
// oracle's X^T W^2 X on the CPU, so that the index bookkeeping of a plan (te() / f() / l() pairs, slices, hash
// splits, key ranges) is checked without a GPU.  The product never loads this library.
#include <algorithm>
#include <vector>

#include "pgb_cells_shared.cuh"

int pgb_model_build_host(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist, int32_t link,
                         double levels, double expectile, int32_t device, const pgb_debug_opts* dbg, pgb_model** out);
CellsView pgb_cells_view_host(const pgb_model* M);

extern "C" int pgb_selfcheck_cells_gram(const pgb_term* terms, int32_t n_terms, int32_t n_features, int32_t dist,
                                        int32_t link, double levels, double expectile, const pgb_debug_opts* opts,
                                        int32_t mode, const double* X, int64_t n, int64_t ld, const double* y,
                                        const float* w, const double* beta, double* out, int32_t* n_generic_pairs) {
  pgb_model* M = nullptr;
  int rc = pgb_model_build_host(terms, n_terms, n_features, dist, link, levels, expectile, 0, opts, &M);
  if (rc != PGB_OK) return rc;
  const pgb_model::Cells& C = M->cells;
  if (!C.enabled) { delete M; return PGB_EUNSUPPORTED; }
  if (n_generic_pairs) *n_generic_pairs = (int32_t)C.gpairs.size();
  const int m = M->m, TR = C.tile_rows;
  std::vector<double> om(n), omz(n), tq((size_t)C.ncol * n, 0.0), sums((size_t)(C.sum_doubles + C.gsum_doubles), 0.0);
  std::vector<uint8_t> cq((size_t)C.ncol * n, 0);
  std::vector<double> b0(m, 0.0);
  double extra[2] = {0.0, 0.0};
  for (int64_t i = 0; i < n; ++i) {
    const double lp = prepass_row_lp(
        M->dterms.data(), n_terms, mode == PGB_MODE_IRLS ? beta : b0.data(), [&](int f) { return X[(int64_t)f * ld + i]; },
        [&](int c, double t) { tq[(size_t)c * n + i] = t; }, [&](int c, int b) { cq[(size_t)c * n + i] = (uint8_t)b; });
    if (mode == PGB_MODE_IRLS) {
      const IrlsRow ir = irls_row(M->fam, lp, y[i], weight_inv(w, i));
      om[i] = ir.omega;
      omz[i] = ir.omega * ir.z;
    } else {
      om[i] = 1.0;
      omz[i] = init_response(M->fam, y[i]);
    }
    extra[0] += om[i];
    extra[1] += omz[i];
  }
  // spline pairs (the arithmetic of cell_rows_pair / cell_rows_diag)
  for (const DCellPair& P : C.pairs) {
    const DTerm& Ti = M->dterms[P.ti];
    const uint32_t hmul = ((uint32_t)P.ncb << 16) / (uint32_t)TR;
    for (int64_t i = 0; i < n; ++i) {
      double a[4], b[4];
      record_powers(Ti.marg[0].nint, tq[(size_t)Ti.fc0 * n + i], a);
      const int ci = cq[(size_t)Ti.fc0 * n + i] & 0x7f;
      if (P.tj >= 0) {
        const DTerm& Tj = M->dterms[P.tj];
        record_powers(Tj.marg[0].nint, tq[(size_t)Tj.fc0 * n + i], b);
        const int cell = ci * P.ncb + (cq[(size_t)Tj.fc0 * n + i] & 0x7f);
        for (int p = 0; p < 4; ++p)
          for (int q = 0; q < 4; ++q) sums[P.sum_off + (int64_t)(4 * p + q) * P.ncells + cell] += om[i] * a[p] * b[q];
      } else {
        const int cell = ci * P.ncb + (int)((((uint32_t)(i % TR)) * hmul) >> 16);
        int e = 0;
        for (int p = 0; p < 4; ++p)
          for (int q = p; q < 4; ++q) sums[P.sum_off + (int64_t)(e++) * P.ncells + cell] += om[i] * a[p] * a[q];
        for (int p = 0; p < 4; ++p) {
          sums[P.sum_off + (int64_t)(10 + p) * P.ncells + cell] += om[i] * a[p];
          sums[P.sum_off + (int64_t)(14 + p) * P.ncells + cell] += omz[i] * a[p];
        }
      }
    }
    for (int k = 0; k < P.ncells; ++k) spline_transform_cell(P.tj < 0, sums.data() + P.sum_off + k, P.ncells);
  }
  // generic pairs: thread (cell, slice) <-> sums[(e + n_acc * slice) * ncells + cell]
  for (const DGenPair& P : C.gpairs) {
    const DGenOrder& O = C.gorders[P.order];
    for (int64_t i = 0; i < n; ++i) {
      const uint32_t key = gen_key(O, (int)(i % C.gtile_rows), [&](int c) { return (uint32_t)cq[(size_t)c * n + i]; });
      if ((key & 0xffffu) == 0xffffu) continue;
      if ((int)key >= P.ncells) { delete M; return PGB_EINVAL; }
      for (int sg = 0; sg < P.nsl; ++sg) {
        double acc[16] = {0};
        gen_row_acc(P, sg, P.wz ? omz[i] : om[i], [&](int c) { return tq[(size_t)c * n + i]; }, acc);
        for (int e = 0; e < P.n_acc; ++e) sums[P.sum_off + (int64_t)(e + P.n_acc * sg) * P.ncells + key] += acc[e];
      }
    }
    for (int ax = 0; ax < P.n_ax; ++ax) {
      const int64_t lines = (int64_t)P.ncells * (P.nprod / 4), sh = (int64_t)1 << (2 * ax);
      for (int64_t idx = 0; idx < lines; ++idx) {
        const int64_t cell = idx % P.ncells, l = idx / P.ncells;
        moment_line(sums.data() + P.sum_off + ((l / sh) * (4 * sh) + (l % sh)) * P.ncells + cell, sh * P.ncells);
      }
    }
  }
  const CellsView V = pgb_cells_view_host(M);
  for (int r = 0; r < m; ++r)
    for (int c = r; c <= m; ++c) {
      double v = 0.0;
      cells_entry_visit(V, r, c, [&](int64_t off, int cnt) {
        if (off < 0) v += extra[off == -1 ? 0 : 1];
        else for (int h = 0; h < cnt; ++h) v += sums[off + h];
      });
      if (c == m) out[(int64_t)m * m + r] = v;
      else out[(int64_t)r * m + c] = out[(int64_t)c * m + r] = v;
    }
  delete M;
  return PGB_OK;
}
