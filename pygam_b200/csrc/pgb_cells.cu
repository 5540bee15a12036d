// pgb_cells.cu -- the Gram pass of models made of plain cubic s() terms (+ intercept): cell-owned
// register accumulation.
//
// Same result as the general path of pgb_gram.cu (G = X^T W^2 X, r = X^T W^2 z; the reference's
// W.B and pseudo data, pygam.py:764-783, before its QR), different decomposition.  For a pair of
// spline terms (i, j) a data row lands in exactly one CELL = (knot interval of x_i, knot interval
// of x_j) and adds the 4 x 4 outer product of its non-zero basis values to the 16 entries of G
// that belong to the cell.  The general path scatters those products into shared memory (16 B of
// read-modify-write per fp64 FMA: 7.7 FMA/clk/SM).  Here a THREAD OWNS A CELL for the whole
// kernel: the rows of a tile are counting-sorted by cell in shared memory, every thread walks the
// rows of its own cell, re-evaluates the two bases from x in registers and accumulates its 16
// sums in registers -- no accumulator traffic at all; the bound is the fp64 pipe.
//
//   kernel 1  gram_prepass_kernel   per row: lp -> mu -> W -> mask -> z (pygam.py:764-777), writes
//                                    omega = W^2 and omega * z (16 B per row) + the log scalars
//   kernel 2  gram_cells_kernel     one ROLE per CTA at a time: a pair (i, j), i < j (16 sums per
//                                    cell), or a term's diagonal block (cells = interval x a hash
//                                    split of the rows; 10 sums + the intercept column + the rhs)
//   kernel 3  cells_slot_reduce     sums the CTA partials of every role in a fixed order
//   kernel 4  cells_assemble        gathers [G | r] from the cell sums (every entry of G collects
//                                    the <= 16 (cell, product) combinations that touch it)
//
// Scheduling: rows are cut into BANDS that fit the L2; inside a band the (role, tile) units are
// dealt role-major to the CTAs in equal shares, so all CTAs work on the same band at the same time
// (the columns come from HBM once per band, the other roles hit the L2) and a CTA serves the same
// one or two roles in every band (its register accumulators are parked in its partial slot when
// it switches).
#include <cuda_runtime.h>
#include <stdio.h>

#include <algorithm>

#include "pgb_internal.h"
#include "pgb_kernels.cuh"
#include "pgb_cells_shared.cuh"

static const int kCellAcc = 18;           // sums per cell (16 for a pair; 10 + 4 + 4 for a diagonal role)
static const int kPreThreads = 256;
static const int kGenAcc = 64;            // sums per thread of the generic kernel (4 x 4 inner x 4 exponents of the first outer factor)
static const int kGenSumThreads = 256;    // CTA of the generic kernel
static const int kRingStages = 2;         // tile buffers per group of the ring kernel
static size_t ring_stage_bytes(int TR, int threads) { return (((size_t)TR * 26 + (size_t)(threads + 8) * 2) + 15) & ~(size_t)15; }
static const int kGenOcc = 1;             // its CTAs per SM (64 sums per thread: about 200 registers)
static const int64_t kBandBytes = 32ll << 20;  // column bytes of a band of rows (stays in the L2)

// tile: t_i, t_j (or omega z), omega; sorted row index; the two interval bytes (later the 16-bit cell key) per row;
// 16-bit counters per (warp, cell)
// counters: per cell one record of W words = 2 W 16-bit counters, one per warp (W = 4 up to 8 warps, else 8)
static int cells_counter_words(int threads) { return threads / 32 <= 8 ? 4 : 8; }
static size_t cells_smem(int tile, int threads) { return (size_t)tile * (24 + 2 + 2) + (size_t)threads * cells_counter_words(threads) * 4 + 64; }

// ----------------------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------------------
// CTAs [c_lo, c_hi] that can serve role r when nr roles are dealt to `grid` CTAs (UnitIter)
static void role_cta_range(int64_t nr, int grid, int r, int& lo, int& hi) {
  lo = grid;
  hi = -1;
  for (int c = 0; c < grid; ++c) {
    const int64_t a = nr * c / grid, b = (nr * (c + 1) + grid - 1) / grid - 1;
    if (r >= a && r <= b) { lo = std::min(lo, c); hi = std::max(hi, c); }
  }
}

// class of a term for the cell-owned pass: 0 intercept, 1 plain cubic s(), 2 generic (te() of two marginals, each a
// plain cubic spline or -- at most one -- an l(); a cubic s() with a by-variable, which is the same design as
// te(s, l); f(); l()), -1: not representable (the model takes the general path of pgb_gram.cu)
static int cells_term_class(const DTerm& t) {
  if (t.kind == PGB_TERM_INTERCEPT) return 0;
  auto cubic_ok = [](const DMarg& g) { return g.order == 3 && !g.periodic && g.nint >= 1 && g.nint <= 127; };
  if (t.kind == PGB_TERM_SPLINE && t.by >= 0) return cubic_ok(t.marg[0]) ? 2 : -1;
  if (t.by >= 0) return -1;
  if (t.kind == PGB_TERM_SPLINE) return (t.fast && cubic_ok(t.marg[0])) ? 1 : -1;
  if (t.kind == PGB_TERM_TENSOR) {
    if (t.n_marg != 2) return -1;
    const bool l0 = t.marg[0].order == PGB_MARG_LINEAR, l1 = t.marg[1].order == PGB_MARG_LINEAR;
    return ((l0 || cubic_ok(t.marg[0])) && (l1 || cubic_ok(t.marg[1])) && !(l0 && l1)) ? 2 : -1;
  }
  if (t.kind == PGB_TERM_FACTOR) return (t.marg[0].order == 0 && !t.marg[0].periodic && t.n_coefs >= 1 && t.n_coefs <= 200) ? 2 : -1;
  if (t.kind == PGB_TERM_LINEAR) return 2;
  return -1;
}

// the (outer cubic factors, inner cubic factors, value factors) shapes gram_cells_generic_kernel instantiates
static bool cells_shape_supported(int n_cubic, int n_val) {
  const int nout = n_cubic > 2 ? n_cubic - 2 : 0, nin = n_cubic < 2 ? n_cubic : 2;
  const int shape = nout * 9 + nin * 3 + n_val;
  switch (shape) {
    case 1 * 9 + 2 * 3 + 0: case 2 * 9 + 2 * 3 + 0: case 0 * 9 + 2 * 3 + 0: case 0 * 9 + 2 * 3 + 1: case 0 * 9 + 1 * 3 + 0:
    case 0 * 9 + 1 * 3 + 1: case 0: case 1: case 2: case 0 * 9 + 2 * 3 + 2: case 1 * 9 + 2 * 3 + 1: case 0 * 9 + 1 * 3 + 2:
      return true;
    default:
      return false;
  }
}

struct HostFactor { int kind, col, n, nspl, stride; };
static void term_factors(const DTerm& t, int cls, std::vector<HostFactor>& out) {
  out.clear();
  if (cls == 1) out.push_back({PGB_FK_CUBIC, t.fc0, t.marg[0].nint, t.marg[0].nspl, 1});
  else if (cls == 2 && t.kind == PGB_TERM_TENSOR) {
    for (int q = 0; q < 2; ++q) {  // first marginal slowest in the coefficient index (an l() marginal has one column)
      const DMarg& g = t.marg[q];
      if (g.order == PGB_MARG_LINEAR) out.push_back({PGB_FK_VALUE, t.fc0 + q, 1, 1, 1});
      else out.push_back({PGB_FK_CUBIC, t.fc0 + q, g.nint, g.nspl, q == 0 ? t.marg[1].nspl : 1});
    }
  } else if (cls == 2 && t.kind == PGB_TERM_SPLINE) {  // s(j, by=k): the basis row times x_k
    out.push_back({PGB_FK_CUBIC, t.fc0, t.marg[0].nint, t.marg[0].nspl, 1});
    out.push_back({PGB_FK_VALUE, t.fc0 + 1, 1, 1, 1});
  } else if (cls == 2 && t.kind == PGB_TERM_FACTOR) out.push_back({PGB_FK_LEVEL, t.fc0, t.n_coefs, t.n_coefs, 1});
  else if (cls == 2 && t.kind == PGB_TERM_LINEAR) out.push_back({PGB_FK_VALUE, t.fc0, 1, 1, 1});
}

int pgb_plan_cells(pgb_model* M) {
  pgb_model::Cells& C = M->cells;
  C = pgb_model::Cells();
  if (M->dbg.force_general) return PGB_OK;  // test-only entry point pgb_debug_model_create
  const int nt = M->n_terms;
  int n_spl = 0, n_gen = 0, icpt = -1, ncol = 0;
  std::vector<int> cls(nt, -1);
  for (int j = 0; j < nt; ++j) {
    DTerm& t = M->dterms[j];
    t.fc0 = -1;
    cls[j] = cells_term_class(t);
    if (cls[j] < 0) return PGB_OK;  // some other term: the general path
    if (cls[j] == 0) {
      if (icpt >= 0) return PGB_OK;
      icpt = j;
      continue;
    }
    t.fc0 = ncol;
    ncol += (t.kind == PGB_TERM_TENSOR || (t.kind == PGB_TERM_SPLINE && t.by >= 0)) ? 2 : 1;
    if (cls[j] == 1) ++n_spl; else ++n_gen;
  }
  if (n_spl + n_gen < 1) return PGB_OK;
  C.icpt_coef = icpt >= 0 ? M->dterms[icpt].coef_offset : -1;
  C.n_spl = n_spl;
  C.ncol = ncol;
  C.mixed = n_gen > 0;
  // CTA shape: one thread per cell of a role.  Pairs of coarse bases (the default n_splines = 20 has 17 x 17
  // = 289 cells) get narrower CTAs, two per SM, so that no thread idles and one CTA sorts while the other sums
  int max_cells = 1;
  for (int i = 0; i < nt; ++i)
    for (int j = i + 1; j < nt; ++j)
      if (cls[i] == 1 && cls[j] == 1)
        max_cells = std::max(max_cells, M->dterms[i].marg[0].nint * M->dterms[j].marg[0].nint);
  if (n_spl <= 1) max_cells = 256;  // diagonal and generic roles only
  C.threads = max_cells <= 256 ? 256 : max_cells <= 320 ? 320 : max_cells <= 384 ? 384 : 512;
  C.occ = C.threads <= 320 ? 2 : 1;  // CTAs per SM of the sort kernels (gram_cells_kernel)
  if (const int t = M->dbg.cells_threads)
    if (t == 256 || t == 320 || t == 384 || t == 512) { C.threads = t; C.occ = t <= 320 ? 2 : 1; }
  // the ring kernel of a prepared shard runs G independent groups of `threads` consumers per CTA, one CTA per SM
  // (0: the barrier version gram_cells_sorted_kernel, two CTAs per SM -- measured faster for the narrow CTAs of coarse
  // bases, whose ring tiles would be half as long)
  C.ring_groups = C.threads <= 384 ? 0 : 1;
  if (M->dbg.cells_kernel == 1) C.ring_groups = 0;
  if (M->dbg.cells_kernel == 2) C.ring_groups = C.threads <= 384 ? 2 : 1;
  const int TH = C.threads;
  int64_t sum_off = 0;
  int64_t n_sub = 0;
  for (int i = 0; i < nt; ++i) {
    const DTerm& ti = M->dterms[i];
    if (cls[i] != 1) continue;
    for (int j = i; j < nt; ++j) {
      const DTerm& tj = M->dterms[j];
      if (cls[j] != 1) continue;
      DCellPair P{};
      P.ti = i;
      P.tj = (j == i) ? -1 : j;
      P.off_i = ti.coef_offset;
      P.off_j = tj.coef_offset;
      P.nint_i = ti.marg[0].nint;
      // diagonal role: the rows of an interval are spread over `ncb` cells by row index, for balance
      P.ncb = (j == i) ? std::max(1, std::min(32, TH / P.nint_i)) : tj.marg[0].nint;
      P.ncells = P.nint_i * P.ncb;
      P.sum_off = sum_off;
      sum_off += (int64_t)kCellAcc * P.ncells;
      const int pair = (int)C.pairs.size();
      C.pairs.push_back(P);
      for (int k0 = 0; k0 < P.ncells; k0 += TH) {
        DCellRole R{};
        R.a = ti.marg[0];
        R.b = tj.marg[0];
        R.diag = (j == i) ? 1 : 0;
        R.ncb = P.ncb;
        R.key0 = k0;
        R.ncell = std::min(TH, P.ncells - k0);
        R.pair = pair;
        R.sa = ti.fc0;
        R.sb = tj.fc0;
        C.roles.push_back(R);
        ++n_sub;
      }
    }
  }
  // a pair split over many roles re-reads its rows once per role: leave very fine bases to the general path
  if ((n_sub > 4 * (int64_t)C.pairs.size() && !C.pairs.empty()) || C.roles.size() > 4096) {
    C = pgb_model::Cells();
    return PGB_OK;
  }
  C.sum_doubles = sum_off;
  // ---- generic pairs: every block with a te() / f() / l() term on one side (incl. its intercept and rhs columns)
  C.term_class.assign(nt, 0);
  for (int j = 0; j < nt; ++j) C.term_class[j] = cls[j];
  C.gpair_of.assign((size_t)(nt + 1) * (nt + 1), -1);
  if (C.mixed) {
    std::vector<HostFactor> fa, fb;
    int64_t gsum_off = 0, starts_off = 0;
    for (int i = 0; i < nt; ++i) {
      for (int j = i; j <= nt; ++j) {
        const int ci = cls[i], cj = (j < nt) ? cls[j] : 0;
        if (ci != 2 && cj != 2) continue;
        term_factors(M->dterms[i], ci, fa);
        if (j < nt) term_factors(M->dterms[j], cj, fb); else fb.clear();
        if (fa.size() + fb.size() > 4) { C = pgb_model::Cells(); return PGB_OK; }
        DGenPair P{};
        P.ta = i;
        P.tb = j;
        P.off_a = M->dterms[i].coef_offset;
        P.off_b = (j < nt) ? M->dterms[j].coef_offset : 0;
        P.wz = (j == nt) ? 1 : 0;
        P.in_a = P.in_b = P.out0 = P.out1 = -1;
        std::vector<int> cubics;
        DGenOrder O{};
        for (int side = 0; side < 2; ++side)
          for (const HostFactor& f : (side ? fb : fa)) {
            const int q = P.nf++;
            P.f_kind[q] = f.kind; P.f_side[q] = side; P.f_col[q] = f.col; P.f_n[q] = f.n;
            P.f_nspl[q] = f.nspl; P.f_stride[q] = f.stride; P.f_key[q] = -1; P.f_ax[q] = -1;
            if (f.kind == PGB_FK_VALUE) { P.val_col[P.nval++] = f.col; continue; }
            if (f.kind == PGB_FK_CUBIC) cubics.push_back(q);
            int kd = -1;
            for (int d = 0; d < O.nkey; ++d) if (O.key_col[d] == f.col) kd = d;
            if (kd < 0) { kd = O.nkey++; O.key_col[kd] = f.col; O.key_n[kd] = f.n; O.key_kind[kd] = f.kind; }
            P.f_key[q] = kd;
          }
        const int nc = (int)cubics.size();
        if (!cells_shape_supported(nc, P.nval)) { C = pgb_model::Cells(); return PGB_OK; }  // general path
        if (nc >= 1) { P.in_b = cubics[nc - 1]; P.f_ax[P.in_b] = 0; }
        if (nc >= 2) { P.in_a = cubics[nc - 2]; P.f_ax[P.in_a] = 1; }
        if (nc >= 3) { P.out0 = cubics[nc - 3]; P.f_ax[P.out0] = 2; }
        if (nc >= 4) { P.out1 = cubics[nc - 4]; P.f_ax[P.out1] = 3; }
        P.n_ax = nc;
        P.n_acc = nc >= 2 ? 16 : nc == 1 ? 4 : 1;
        P.nsl = nc >= 4 ? 16 : nc == 3 ? 4 : 1;
        P.nprod = P.n_acc * P.nsl;
        P.ne = nc >= 3 ? 4 : 1;
        P.nslt = P.nsl / P.ne;
        int64_t kc = 1;
        for (int d = O.nkey - 1; d >= 0; --d) { O.key_mul[d] = (int32_t)kc; kc *= O.key_n[d]; }
        // hash split of the rows of a cell: runs of about 32 rows per (cell, split) and generic tile
        O.nh = (int)std::max<int64_t>(1, std::min<int64_t>(256, 1024 / kc));
        if (kc * O.nh > 65000) { C = pgb_model::Cells(); return PGB_OK; }
        O.ncells = (int32_t)(kc * O.nh);
        O.spitch = (O.ncells + 2 + 7) & ~7;
        int oi = -1;
        for (size_t q = 0; q < C.gorders.size() && oi < 0; ++q) {
          const DGenOrder& Q = C.gorders[q];
          bool same = Q.nkey == O.nkey && Q.nh == O.nh;
          for (int d = 0; d < O.nkey && same; ++d) same = Q.key_col[d] == O.key_col[d] && Q.key_n[d] == O.key_n[d];
          if (same) oi = (int)q;
        }
        if (oi < 0) {
          O.starts_off = starts_off;
          starts_off += O.spitch;
          oi = (int)C.gorders.size();
          C.gorders.push_back(O);
        }
        const DGenOrder& Q = C.gorders[oi];
        P.order = oi; P.nh = Q.nh; P.ncells = Q.ncells; P.spitch = Q.spitch; P.starts_off = Q.starts_off;
        for (int d = 0; d < 4; ++d) P.key_mul[d] = Q.key_mul[d];
        P.sum_off = C.sum_doubles + gsum_off;
        gsum_off += (int64_t)P.nprod * P.ncells;
        const int pi = (int)C.gpairs.size();
        C.gpair_of[(size_t)i * (nt + 1) + j] = pi;
        C.gpairs.push_back(P);
        const int ncl = std::max(1, kGenSumThreads / P.nslt);
        for (int k0 = 0; k0 < P.ncells; k0 += ncl) {
          DGenRole R{};
          R.pair = pi; R.key0 = k0; R.ncl = std::min(ncl, P.ncells - k0);
          C.groles.push_back(R);
        }
      }
    }
    C.gsum_doubles = gsum_off;
    C.gstarts_u16 = starts_off;
    if (C.groles.size() > 16384 || C.gorders.size() > 4096) { C = pgb_model::Cells(); return PGB_OK; }
  }
  // units are dealt to kSMs x G "virtual CTAs": the groups of the ring kernel, the CTAs of the sort kernels
  C.grid = kSMs * (C.ring_groups > 0 ? C.ring_groups : 2);
  C.ggrid = kSMs * kGenOcc;
  C.tile_rows = 8192;
  if (M->dbg.cells_tile_rows > 0) C.tile_rows = std::max(64, M->dbg.cells_tile_rows & ~63);
  // 1 KB per CTA for the kernels' static shared memory and the driver's reservation
  while (cells_smem(C.tile_rows, TH) + 1024 > (C.occ == 2 ? kSmemMax / 2 - 1024 : kSmemMax)) C.tile_rows -= 64;
  if (C.ring_groups > 0)
    while ((size_t)C.ring_groups * kRingStages * ring_stage_bytes(C.tile_rows, TH) + 1024 > kSmemMax) C.tile_rows -= 64;
  else
    while ((size_t)C.tile_rows * 26 + (TH + 8) * 2 + 1024 > kSmemMax / 2 - 1024) C.tile_rows -= 64;
  for (DCellRole& R : C.roles) R.hmul = ((uint32_t)R.ncb << 16) / (uint32_t)C.tile_rows;
  // The generic orders sort much longer tiles than the spline roles: a thread's run of rows must be long enough
  // to amortise its start (three dependent gathers) and to even out the lanes of a warp; nothing of a generic
  // tile is staged in shared memory, so only the 16-bit row index bounds it
  C.gtile_rows = M->dbg.cells_tile_rows > 0 ? 8 * C.tile_rows : 32768;
  for (DGenOrder& O : C.gorders) O.hmul = ((uint32_t)O.nh << 16) / (uint32_t)C.gtile_rows;
  C.smem = cells_smem(C.tile_rows, TH);
  C.slots_per_cta = ((int)C.roles.size() + C.grid - 1) / C.grid + 1;
  C.gslots_per_cta = 0;
  if (!C.groles.empty()) {
    // weight of a role's tile: the dependent gathers of its fullest threads, about rows per cell + a constant
    const int nr = (int)C.groles.size();
    C.grole_prefix.assign(nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
      const DGenPair& P = C.gpairs[C.groles[r].pair];
      const int w = 3 + (int)std::min<int64_t>(250, ((int64_t)C.gtile_rows + P.ncells - 1) / P.ncells);
      C.grole_prefix[r + 1] = C.grole_prefix[r] + w;
    }
    const int64_t W = C.grole_prefix[nr];
    C.gcta_role_lo.assign(C.ggrid, 0);
    for (int r = 0; r < nr; ++r) { C.groles[r].c_lo = C.ggrid; C.groles[r].c_hi = -1; }
    for (int c = 0, r0 = 0; c < C.ggrid; ++c) {
      // CTA c covers [W c / grid, W (c + 1) / grid) of the weight axis; role r covers [prefix[r], prefix[r + 1])
      while (r0 < nr - 1 && (int64_t)C.grole_prefix[r0 + 1] * C.ggrid <= W * c) ++r0;
      C.gcta_role_lo[c] = r0;
      int cnt = 0;
      for (int r = r0; r < nr && (int64_t)C.grole_prefix[r] * C.ggrid < W * (c + 1); ++r) {
        C.groles[r].c_lo = std::min(C.groles[r].c_lo, c);
        C.groles[r].c_hi = std::max(C.groles[r].c_hi, c);
        ++cnt;
      }
      C.gslots_per_cta = std::max(C.gslots_per_cta, cnt + 1);
    }
  }
  C.slot_doubles = (int64_t)kCellAcc * TH;
  C.gslot_doubles = (int64_t)kGenAcc * kGenSumThreads;
  for (int r = 0; r < (int)C.roles.size(); ++r) role_cta_range((int64_t)C.roles.size(), C.grid, r, C.roles[r].c_lo, C.roles[r].c_hi);
  // coefficient -> (term, local index) for the assembly
  C.coef_term.assign(M->m, -1);
  for (int j = 0; j < nt; ++j)
    for (int q = 0; q < M->dterms[j].n_coefs; ++q) C.coef_term[M->dterms[j].coef_offset + q] = j;
  // (term i, term j) -> pair
  C.pair_of.assign((size_t)nt * nt, -1);
  for (size_t p = 0; p < C.pairs.size(); ++p) {
    const DCellPair& P = C.pairs[p];
    C.pair_of[(size_t)P.ti * nt + (P.tj < 0 ? P.ti : P.tj)] = (int)p;
  }
  C.enabled = true;
  return PGB_OK;
}

int pgb_cells_upload(pgb_model* M) {
  pgb_model::Cells& C = M->cells;
  if (!C.enabled) return PGB_OK;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dptr, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dptr, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess && bytes) e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&C.d_roles, C.roles.data(), C.roles.size() * sizeof(DCellRole));
  up((void**)&C.d_pairs, C.pairs.data(), C.pairs.size() * sizeof(DCellPair));
  up((void**)&C.d_coef_term, C.coef_term.data(), C.coef_term.size() * sizeof(int32_t));
  up((void**)&C.d_pair_of, C.pair_of.data(), C.pair_of.size() * sizeof(int32_t));
  up((void**)&C.d_gorders, C.gorders.data(), C.gorders.size() * sizeof(DGenOrder));
  up((void**)&C.d_gpairs, C.gpairs.data(), C.gpairs.size() * sizeof(DGenPair));
  up((void**)&C.d_groles, C.groles.data(), C.groles.size() * sizeof(DGenRole));
  up((void**)&C.d_term_class, C.term_class.data(), C.term_class.size() * sizeof(int32_t));
  up((void**)&C.d_gpair_of, C.gpair_of.data(), C.gpair_of.size() * sizeof(int32_t));
  up((void**)&C.d_grole_prefix, C.grole_prefix.data(), C.grole_prefix.size() * sizeof(int32_t));
  up((void**)&C.d_gcta_role_lo, C.gcta_role_lo.data(), C.gcta_role_lo.size() * sizeof(int32_t));
  if (e != cudaSuccess) return pgb_cuda_fail(e, "pgb_cells_upload");
  return PGB_OK;
}

void pgb_cells_release(pgb_model* M) {
  pgb_model::Cells& C = M->cells;
  cudaFree(C.d_roles);
  cudaFree(C.d_pairs);
  cudaFree(C.d_coef_term);
  cudaFree(C.d_pair_of);
  cudaFree(C.d_gorders);
  cudaFree(C.d_gpairs);
  cudaFree(C.d_groles);
  cudaFree(C.d_term_class);
  cudaFree(C.d_gpair_of);
  cudaFree(C.d_grole_prefix);
  cudaFree(C.d_gcta_role_lo);
  C.d_grole_prefix = nullptr; C.d_gcta_role_lo = nullptr;
  C.d_roles = nullptr; C.d_pairs = nullptr; C.d_coef_term = nullptr; C.d_pair_of = nullptr;
  C.d_gorders = nullptr; C.d_gpairs = nullptr; C.d_groles = nullptr; C.d_term_class = nullptr; C.d_gpair_of = nullptr;
}

CellsView pgb_cells_view_host(const pgb_model* M) {
  const pgb_model::Cells& C = M->cells;
  return CellsView{C.coef_term.data(), C.pair_of.data(), C.term_class.data(), C.gpair_of.data(), C.pairs.data(),
                   C.gpairs.data(), M->n_terms, M->m, C.icpt_coef};
}

// host-only self check: every structurally non-zero entry of the upper triangle of [G | r] must be assembled
// from at least one (pair, product, cell) combination, and every combination of every pair must be used
// (exactly once for the spline pairs; a symmetric generic block holds both triangles, of which the assembly
// reads one); returns the number of violations
int pgb_cells_coverage_errors(const pgb_model* M) {
  const pgb_model::Cells& C = M->cells;
  if (!C.enabled) return 0;
  const int m = M->m, nt = M->n_terms;
  int errors = 0;
  std::vector<int> used((size_t)(C.sum_doubles + C.gsum_doubles), 0);
  const CellsView V = pgb_cells_view_host(M);
  for (int r = 0; r < m; ++r)
    for (int c = r; c <= m; ++c) {
      int hits = 0;
      cells_entry_visit(V, r, c, [&](int64_t off, int cnt) {
        ++hits;
        for (int h = 0; h < cnt; ++h) ++used[off + h];
      });
      const int tr = C.coef_term[r], tc = (c < m) ? C.coef_term[c] : nt;
      bool structural_zero = false;  // same term, coefficients further apart than the support of a cubic / another level
      if (tr == tc) {
        const DTerm& T = M->dterms[tr];
        const int la = r - T.coef_offset, lb = c - T.coef_offset;
        if (T.kind == PGB_TERM_SPLINE) structural_zero = lb - la > 3;
        else if (T.kind == PGB_TERM_FACTOR) structural_zero = la != lb;
        else if (T.kind == PGB_TERM_TENSOR) {
          const int K1 = T.marg[1].nspl;
          structural_zero = std::abs(la / K1 - lb / K1) > 3 || std::abs(la % K1 - lb % K1) > 3;
        }
      }
      if (hits < 1 && !structural_zero) ++errors;
      if (hits >= 1 && structural_zero) ++errors;
    }
  for (const DCellPair& P : C.pairs) {
    const int n_acc = (P.tj < 0) ? 18 : 16;
    for (int e = 0; e < n_acc; ++e)
      for (int k = 0; k < P.ncells; ++k) {
        // without an intercept the diagonal role's intercept-column sums stay unused
        const int want = (P.tj < 0 && e >= 10 && e < 14 && C.icpt_coef < 0) ? 0 : 1;
        if (used[P.sum_off + (int64_t)e * P.ncells + k] != want) ++errors;
      }
  }
  for (const DGenPair& P : C.gpairs) {
    // the moments of cells next to the edge of a marginal are used by fewer entries, but every (product, cell)
    // of an off-diagonal pair feeds exactly one entry; a diagonal pair (ta == tb) holds both triangles
    for (int64_t q = 0; q < (int64_t)P.nprod * P.ncells; ++q) {
      const int u = used[P.sum_off + q];
      if (P.ta != P.tb ? u != 1 : u > 1) ++errors;
    }
  }
  // every cell of every pair belongs to exactly one role
  for (size_t p = 0; p < C.pairs.size(); ++p) {
    int covered = 0;
    for (const DCellRole& R : C.roles)
      if (R.pair == (int)p) covered += R.ncell;
    if (covered != C.pairs[p].ncells) ++errors;
  }
  for (size_t p = 0; p < C.gpairs.size(); ++p) {
    int covered = 0;
    for (const DGenRole& R : C.groles)
      if (R.pair == (int)p) covered += R.ncl;
    if (covered != C.gpairs[p].ncells) ++errors;
  }
  return errors;
}

// workspace: [CTA partial slots: spline roles | generic roles][cell sums: spline pairs | generic pairs][scalar partials]
// [extra][omega][omega z][record coordinate of every factor column][cell bytes][stored order: spline roles | generic orders]
static int64_t rows_pad(int64_t n) { return (n + 15) & ~(int64_t)15; }
extern "C" int64_t pgb_cells_ws_bytes(const pgb_model* M, int64_t n, int with_order) {
  const pgb_model::Cells& C = M->cells;
  const int64_t slot_all = ((int64_t)C.grid * C.slots_per_cta * C.slot_doubles + (int64_t)C.ggrid * C.gslots_per_cta * C.gslot_doubles);
  const int64_t np = rows_pad(n);
  const int64_t ntl = (n + C.tile_rows - 1) / C.tile_rows;
  const int64_t gntl = C.mixed ? (n + C.gtile_rows - 1) / C.gtile_rows : 0;
  // sorted rows + cell starts (pgb_gram_prepare); not needed by passes that sort inside the kernel.  A model with
  // generic pairs always sorts into the workspace (its streaming pass runs the sort kernels first)
  int64_t order = 0;
  if (with_order || C.mixed)
    order = ((int64_t)C.roles.size() * ntl * (C.tile_rows + C.threads + 8) + (int64_t)C.gorders.size() * gntl * C.gtile_rows +
             C.gstarts_u16 * gntl) * 2 + 512;
  return (slot_all + C.sum_doubles + C.gsum_doubles + (int64_t)kSMs * 8 * (PGB_GS_COUNT + 2) + 8 + (2 + C.ncol) * np) * 8 +
         (int64_t)C.ncol * np + order + 2048;
}

// ----------------------------------------------------------------------------------------
// kernel 1: one thread per row.  Basis of every spline term -> lp -> the IRLS step (mu, W, mask,
// working response; pygam.py:764-777) or the response of _initial_estimate; writes per row omega
// and omega * z, and per (row, spline term) the row's CELL RECORD: the knot interval (1 byte, bit 7:
// the row lies outside the edge knots) and the local coordinate t in [0, 1] inside it (fp64; for a
// row outside the knots the normalised coordinate u itself) -- 9 bytes instead of four (value,
// index) pairs: the cells kernel re-evaluates the four cubic B-spline values from t in registers.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void extrap_basis(const DMarg& mg, double u, double* v) { extrap_basis_n(mg.nint, u, v); }

struct PreSpl {  // what the prepass needs of one spline term
  double lo, inv_span, nint_d;
  int32_t feature, nint, coef_offset, term;
};

__global__ void __launch_bounds__(kPreThreads, 3)
gram_prepass_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam, int mode,
                    const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                    const float* __restrict__ w, const double* __restrict__ beta, int64_t np, double* __restrict__ om,
                    double* __restrict__ omz, double* __restrict__ tq, uint8_t* __restrict__ cq,
                    double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + (PGB_GS_COUNT + 2) * 32);
  PreSpl* s_spl = reinterpret_cast<PreSpl*>(s_terms + n_terms);
  __shared__ int s_nspl, s_icpt;
  for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();
  if (tid == 0) {
    int ns = 0, ic = -1;
    for (int q = 0; q < n_terms; ++q) {
      const DTerm& T = s_terms[q];
      if (T.kind == PGB_TERM_INTERCEPT) { ic = T.coef_offset; continue; }
      const DMarg& mg = T.marg[0];
      s_spl[ns++] = PreSpl{mg.lo, mg.inv_span, (double)mg.nint, mg.feature, mg.nint, T.coef_offset, q};
    }
    s_nspl = ns;
    s_icpt = ic;
  }
  __syncthreads();
  const int n_spl = s_nspl;
  const double icpt_beta = s_icpt >= 0 ? s_beta[s_icpt] : 0.0;
  double sc[PGB_GS_COUNT + 2];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT + 2; ++s) sc[s] = 0.0;
  constexpr int NX = 8;  // feature columns of a row fetched together: 8 loads in flight per thread
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yi = __ldg(y + i);
    const double winv = weight_inv(w, i);
    double lp = icpt_beta;
    for (int base = 0; base < n_spl; base += NX) {
      double xv[NX];
#pragma unroll
      for (int q = 0; q < NX; ++q)
        if (base + q < n_spl) xv[q] = __ldg(X + (int64_t)s_spl[base + q].feature * ld + i);
#pragma unroll
      for (int q = 0; q < NX; ++q) {
        if (base + q >= n_spl) break;
        const PreSpl& S = s_spl[base + q];
        const double u = (xv[q] - S.lo) * S.inv_span;
        double v[4], t;
        int cell;
        if (u >= 0.0 && u <= 1.0) {  // the arithmetic of eval_term's straight-line cubic
          const double s = u * S.nint_d;
          cell = min((int)s, S.nint - 1);
          t = s - (double)cell;
          cubic_bspline(t, v[0], v[1], v[2], v[3]);
          t -= 0.5;  // the record holds the centred coordinate, in [-0.5, 0.5]
        } else {
          extrap_basis(s_terms[S.term].marg[0], u, v);
          cell = (u < 0.0 ? 0 : S.nint - 1) | 0x80;
          t = (u < 0.0) ? u - 1.0 : u + 1.0;  // outside the knots: the normalised coordinate, moved out of [-0.5, 0.5]
        }
        const double* bt = s_beta + S.coef_offset + (cell & 0x7f);
        lp = fma(v[0], bt[0], lp);
        lp = fma(v[1], bt[1], lp);
        lp = fma(v[2], bt[2], lp);
        lp = fma(v[3], bt[3], lp);
        tq[(int64_t)(base + q) * np + i] = t;
        cq[(int64_t)(base + q) * np + i] = (uint8_t)cell;
      }
    }
    double o, oz;
    if (mode == PGB_MODE_IRLS) {
      const IrlsRow ir = irls_row(fam, lp, yi, winv);
      o = ir.omega;
      oz = ir.omega * ir.z;
      if (ir.keep) {
        sc[PGB_GS_NUSED] += 1.0;
        sc[PGB_GS_DEV] += dist_dev(fam, yi, ir.mu);
        sc[PGB_GS_NCORRECT] += (yi == ((ir.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      o = 1.0;
      oz = init_response(fam, yi);
      sc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(oz)) sc[PGB_GS_NONFINITE] += 1.0;
    }
    sc[PGB_GS_NROWS] += 1.0;
    sc[PGB_GS_COUNT] += o;       // G[intercept, intercept]
    sc[PGB_GS_COUNT + 1] += oz;  // r[intercept]
    om[i] = o;
    omz[i] = oz;
  }
  block_reduce_store<PGB_GS_COUNT + 2>(sc, s_red, scal_part + (int64_t)blockIdx.x * (PGB_GS_COUNT + 2));
}

// ----------------------------------------------------------------------------------------
// kernel 2: cells
// ----------------------------------------------------------------------------------------
// The four cubic B-splines of an interval are polynomials of the centred local coordinate s = t - 1/2:
// b_a(s) = sum_p kM[a][p] s^p.  The cells kernel therefore sums MOMENTS omega s_i^p s_j^q (p, q = 0..3) --
// 2 + 2 + 3 multiplies and 16 FMAs per (row, pair) instead of two basis evaluations (24) + 4 + 16 -- and
// cells_transform_kernel maps the 16 moment sums of a cell to the 16 entries of G: B = M S M^T.

// A row outside the edge knots (rare: custom edge_knots narrower than the data) has four basis values v
// that are no cubic of s; its "powers" are mu = M^-1 v, so that M mu = v and the same sums stay valid.
__device__ __noinline__ void cell_powers_outside(const DMarg* mg, double enc, double* mu) {
  const double u = enc < 0.0 ? enc + 1.0 : enc - 1.0;
  double v[PGB_MAXV];
  extrap_basis(*mg, u, v);
  mu[0] = v[0] + v[1] + v[2] + v[3];
  mu[1] = 1.5 * (v[3] - v[0]) + 0.5 * (v[2] - v[1]);
  mu[2] = (23.0 / 12.0) * (v[0] + v[3]) - (1.0 / 12.0) * (v[1] + v[2]);
  mu[3] = 1.875 * (v[3] - v[0]) + 0.375 * (v[1] - v[2]);
}

template <bool CAREFUL>
__device__ __forceinline__ void cell_powers(const DMarg* mg, double s, double& p0, double& p1, double& p2, double& p3) {
  if (CAREFUL && fabs(s) > 0.75) {
    double mu[4];
    cell_powers_outside(mg, s, mu);
    p0 = mu[0]; p1 = mu[1]; p2 = mu[2]; p3 = mu[3];
  } else {
    p0 = 1.0;
    p1 = s;
    p2 = s * s;
    p3 = p2 * s;
  }
}

// one row of a pair cell: acc[4 p + q] += omega s_i^p s_j^q.  Inside the knots the powers are chained through
// omega (u_p = u_{p-1} s_i): 3 + 2 multiplies, 4 adds and 12 FMAs = 21 fp64 instructions for 16 sums.
template <bool CAREFUL>
__device__ __forceinline__ void pair_row(const DCellRole* R, double sa, double sb, double ww, double (&acc)[kCellAcc]) {
  if (CAREFUL) {
    double a0, a1, a2, a3, b0, b1, b2, b3;
    cell_powers<true>(&R->a, sa, a0, a1, a2, a3);
    cell_powers<true>(&R->b, sb, b0, b1, b2, b3);
    a0 *= ww; a1 *= ww; a2 *= ww; a3 *= ww;
    acc[0] = fma(a0, b0, acc[0]); acc[1] = fma(a0, b1, acc[1]); acc[2] = fma(a0, b2, acc[2]); acc[3] = fma(a0, b3, acc[3]);
    acc[4] = fma(a1, b0, acc[4]); acc[5] = fma(a1, b1, acc[5]); acc[6] = fma(a1, b2, acc[6]); acc[7] = fma(a1, b3, acc[7]);
    acc[8] = fma(a2, b0, acc[8]); acc[9] = fma(a2, b1, acc[9]); acc[10] = fma(a2, b2, acc[10]); acc[11] = fma(a2, b3, acc[11]);
    acc[12] = fma(a3, b0, acc[12]); acc[13] = fma(a3, b1, acc[13]); acc[14] = fma(a3, b2, acc[14]); acc[15] = fma(a3, b3, acc[15]);
  } else {
    const double u1 = ww * sa, u2 = u1 * sa, u3 = u2 * sa;
    const double b2 = sb * sb, b3 = b2 * sb;
    acc[0] += ww; acc[1] = fma(ww, sb, acc[1]); acc[2] = fma(ww, b2, acc[2]); acc[3] = fma(ww, b3, acc[3]);
    acc[4] += u1; acc[5] = fma(u1, sb, acc[5]); acc[6] = fma(u1, b2, acc[6]); acc[7] = fma(u1, b3, acc[7]);
    acc[8] += u2; acc[9] = fma(u2, sb, acc[9]); acc[10] = fma(u2, b2, acc[10]); acc[11] = fma(u2, b3, acc[11]);
    acc[12] += u3; acc[13] = fma(u3, sb, acc[13]); acc[14] = fma(u3, b2, acc[14]); acc[15] = fma(u3, b3, acc[15]);
  }
}

// one row of a diagonal cell: the 10 moment sums of the packed upper triangle (p <= q), the intercept column
// (omega s^p), the right-hand side (omega z s^p)
template <bool CAREFUL>
__device__ __forceinline__ void diag_row(const DCellRole* R, double sa, double z, double ww, double (&acc)[kCellAcc]) {
  double a0, a1, a2, a3;
  cell_powers<CAREFUL>(&R->a, sa, a0, a1, a2, a3);
  const double v0 = a0 * ww, v1 = a1 * ww, v2 = a2 * ww, v3 = a3 * ww;
  acc[0] = fma(v0, a0, acc[0]); acc[1] = fma(v0, a1, acc[1]); acc[2] = fma(v0, a2, acc[2]); acc[3] = fma(v0, a3, acc[3]);
  acc[4] = fma(v1, a1, acc[4]); acc[5] = fma(v1, a2, acc[5]); acc[6] = fma(v1, a3, acc[6]);
  acc[7] = fma(v2, a2, acc[7]); acc[8] = fma(v2, a3, acc[8]); acc[9] = fma(v3, a3, acc[9]);
  acc[10] += v0; acc[11] += v1; acc[12] += v2; acc[13] += v3;
  acc[14] = fma(a0, z, acc[14]); acc[15] = fma(a1, z, acc[15]); acc[16] = fma(a2, z, acc[16]); acc[17] = fma(a3, z, acc[17]);
}

// the rows of one cell, in the stored order, two rows in flight (their gathers are issued together)
template <bool CAREFUL>
__device__ __forceinline__ void cell_rows_pair(const DCellRole* R, const uint16_t* __restrict__ perm, int c,
                                               const double* __restrict__ ta, const double* __restrict__ tb,
                                               const double* __restrict__ wv, double (&acc)[kCellAcc]) {
  int s = 0;
#pragma unroll 1
  for (; s + 1 < c; s += 2) {
    const int r0 = perm[s], r1 = perm[s + 1];
    const double w0 = wv[r0], x0 = ta[r0], y0 = tb[r0];
    const double w1 = wv[r1], x1 = ta[r1], y1 = tb[r1];
    pair_row<CAREFUL>(R, x0, y0, w0, acc);
    pair_row<CAREFUL>(R, x1, y1, w1, acc);
  }
  if (s < c) {
    const int r0 = perm[s];
    pair_row<CAREFUL>(R, ta[r0], tb[r0], wv[r0], acc);
  }
}

template <bool CAREFUL>
__device__ __forceinline__ void cell_rows_diag(const DCellRole* R, const uint16_t* __restrict__ perm, int c,
                                               const double* __restrict__ ta, const double* __restrict__ wz,
                                               const double* __restrict__ wv, double (&acc)[kCellAcc]) {
  int s = 0;
#pragma unroll 1
  for (; s + 1 < c; s += 2) {
    const int r0 = perm[s], r1 = perm[s + 1];
    const double w0 = wv[r0], x0 = ta[r0], z0 = wz[r0];
    const double w1 = wv[r1], x1 = ta[r1], z1 = wz[r1];
    diag_row<CAREFUL>(R, x0, z0, w0, acc);
    diag_row<CAREFUL>(R, x1, z1, w1, acc);
  }
  if (s < c) {
    const int r0 = perm[s];
    diag_row<CAREFUL>(R, ta[r0], wz[r0], wv[r0], acc);
  }
}

// the (role, tile) units of one CTA, band by band; odd bands are walked backwards so that a band
// starts with the role the previous one ended with
struct UnitIter {
  int64_t ntiles, nbands, band, uu, u_lo, u_hi, nb, t0;
  int band_tiles, n_roles, cta, ncta;
  __device__ void set_band() {
    t0 = band * band_tiles;
    nb = min((int64_t)band_tiles, ntiles - t0);
    const int64_t U = nb * n_roles;
    u_lo = U * cta / ncta;
    u_hi = U * (cta + 1) / ncta;
    uu = 0;
  }
  __device__ void skip_empty() {
    while (band < nbands && uu >= u_hi - u_lo) {
      ++band;
      if (band < nbands) set_band();
    }
  }
  __device__ void init(int64_t n, int TR, int bt, int nr) { init(n, TR, bt, nr, (int)blockIdx.x, (int)gridDim.x); }
  // (c, nc): the units of virtual CTA c of nc (a CTA of the ring kernel runs several independent groups)
  __device__ void init(int64_t n, int TR, int bt, int nr, int c, int nc) {
    cta = c;
    ncta = nc;
    band_tiles = bt;
    n_roles = nr;
    ntiles = (n + TR - 1) / TR;
    nbands = (ntiles + bt - 1) / bt;
    band = 0;
    set_band();
    skip_empty();
  }
  __device__ bool valid() const { return band < nbands; }
  __device__ int64_t unit() const { return (band & 1) ? (u_hi - 1 - uu) : (u_lo + uu); }
  __device__ int role() const { return (int)(unit() / nb); }
  __device__ int64_t tile() const { return t0 + unit() % nb; }
  __device__ void next() {
    ++uu;
    skip_empty();
  }
};

// Bank-aware order inside every cell (pgb_gram_refine).  The sum kernels gather t_i, t_j and omega of row
// perm[start + s] for the 32 cells of a warp at once: a 64-bit shared-memory load of 32 unrelated rows costs about
// 4.2 wavefronts where 2 is the minimum (both half-warps on 16 distinct bank pairs, bank pair = row & 15) -- the
// shared-memory pipe, not the fp64 pipe, bounded the kernel.  The order of the rows of a cell is free, so step by
// step every lane picks, among the rows its cell has left, one whose bank pair no other lane of its half-warp has
// taken (up to four rounds of compare / re-pick; a lane that runs out of candidates keeps a conflicting row).
// Simulated cost per half-warp gather: 2.20 -> 1.70 wavefronts; measured: cells kernel of config 2 -13 %.
__device__ __forceinline__ void bank_aware_reorder(uint16_t* run, int c, int lane) {
  int maxc = c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
  const unsigned half = (unsigned)lane & 16u;
  for (int s = 0; s < maxc; ++s) {
    const bool act = s < c;
    int cand = s;
    bool settled = false;
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
      // lanes of my half-warp whose candidate sits in my bank pair: five ballots over the bits of the key
      // (the match instruction is an order of magnitude slower)
      const unsigned key = ((unsigned)run[act ? cand : 0] & 15u) | half;
      unsigned same = __ballot_sync(0xffffffffu, act);
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (key >> b) & 1u);
        same &= ((key >> b) & 1u) ? bal : ~bal;
      }
      const unsigned done = __ballot_sync(0xffffffffu, settled) & same;
      const int winner = done ? (__ffs(done) - 1) : (__ffs(same) - 1);
      if (act && !settled) {
        if (lane == winner) settled = true;
        else if (cand + 1 < c) ++cand;
      }
      if (__ballot_sync(0xffffffffu, act && !settled) == 0u) break;  // every lane has its bank pair
    }
    if (act && cand != s) {
      const uint16_t tmp = run[s];
      run[s] = run[cand];
      run[cand] = tmp;
    }
    __syncwarp();
  }
}

// pgb_gram_refine: the stored order of every (role, tile) of a prepared shard, cell by cell, in bank-aware order.
// One CTA per unit: the tile's row indices and cell starts are staged in shared memory, reordered, written back.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
gram_cells_refine_kernel(int64_t n, int TR, int n_roles, uint16_t* __restrict__ perm_g, const uint16_t* __restrict__ starts_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* perm = reinterpret_cast<uint16_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t ntl = (n + TR - 1) / TR;
  for (int64_t u = blockIdx.x; u < ntl * n_roles; u += gridDim.x) {
    const int64_t tile = u % ntl;
    const int cnt = (int)min((int64_t)TR, n - tile * TR);
    uint32_t* pg = reinterpret_cast<uint32_t*>(perm_g + u * TR);
    const uint16_t* sg = starts_g + u * (THREADS + 8);
    __syncthreads();
    for (int i = tid; i < (cnt + 1) / 2; i += THREADS) reinterpret_cast<uint32_t*>(perm)[i] = pg[i];
    const int start = sg[tid], c = (int)sg[tid + 1] - start;
    __syncthreads();
    bank_aware_reorder(perm + start, c, lane);
    __syncthreads();
    for (int i = tid; i < (cnt + 1) / 2; i += THREADS) pg[i] = reinterpret_cast<const uint32_t*>(perm)[i];
  }
}

// EMIT = false: sort the rows of every tile and sum them right away (rows streamed through once: the
// host-buffer path).  EMIT = true: sort only and write the order -- per (role, tile) the sorted row indices and
// the start offset of every cell -- for gram_cells_sorted_kernel: the cells of a row do not depend on beta,
// so a fit sorts its rows once (pgb_gram_prepare) and every PIRLS iteration reuses the order.
template <int THREADS, int MINB, bool EMIT>
__global__ void __launch_bounds__(THREADS, MINB)
gram_cells_kernel(int64_t n, int64_t np, const double* __restrict__ om, const double* __restrict__ omz,
                  const double* __restrict__ tq, const uint8_t* __restrict__ cq, const DCellRole* __restrict__ roles,
                  int n_roles, int TR, int band_tiles, int slots_per_cta, double* __restrict__ part,
                  uint16_t* __restrict__ perm_g, uint16_t* __restrict__ starts_g) {
  constexpr int kCellThreads = THREADS, kCellWarps = THREADS / 32;
  constexpr int W = (kCellWarps <= 8) ? 4 : 8;  // counter words per cell
  constexpr int SH = (W == 8) ? 2 : 3;          // a cell's record starts in bank group (cell & (32 / W - 1))
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ DCellRole R;
  __shared__ int s_wtot[32];
  __shared__ int s_careful;
  __shared__ __align__(8) unsigned long long s_bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* ta = reinterpret_cast<double*>(smem_raw);   // t of the row term
  double* tb = ta + TR;                               // t of the column term (diagonal role: omega z)
  double* wv = tb + TR;                               // omega
  uint16_t* perm = reinterpret_cast<uint16_t*>(wv + TR);
  // 16-bit row counters, one per (cell, warp): cell k's record is W words; warp q's counter is half (q & 1)
  // of word ((q >> 1) + (k >> SH)) & (W - 1) -- the rotation spreads a warp's atomics over all 32 banks
  uint32_t* wh32 = reinterpret_cast<uint32_t*>(perm + TR);
  // interval bytes of the two terms; pass 1 overwrites them with the 16-bit cell keys of the rows (a thread's
  // four rows: keys 0, 1 in its word of ca, keys 2, 3 in its word of cb)
  uint8_t* ca = reinterpret_cast<uint8_t*>(wh32 + kCellThreads * W);
  uint8_t* cb = ca + TR;
  const uint32_t bar_c = smem_addr(&s_bar[0]), bar_b = smem_addr(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar_c, 1);
    mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  const int role_lo = (int)((int64_t)n_roles * blockIdx.x / gridDim.x);  // first role this CTA can ever serve
  double* my_part = part + (int64_t)blockIdx.x * slots_per_cta * (kCellAcc * kCellThreads);
  int cur_role = -1;
  uint32_t phase = 0;
  double acc[kCellAcc];
#pragma unroll
  for (int e = 0; e < kCellAcc; ++e) acc[e] = 0.0;
  UnitIter it;
  it.init(n, TR, band_tiles, n_roles);
  // the interval bytes of a unit: a bulk copy of 2 bytes per row, issued one unit ahead
  auto load_cells = [&](const UnitIter& u) {
    const int role = u.role();
    const int64_t row0 = u.tile() * TR;
    const uint32_t cnt16 = (uint32_t)((min((int64_t)TR, n - row0) + 15) & ~(int64_t)15);
    const int sa = roles[role].sa, sb = roles[role].sb;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses vs the async writes
    mbar_expect_tx(bar_c, 2u * cnt16);
    bulk_g2s(smem_addr(ca), cq + (int64_t)sa * np + row0, cnt16, bar_c);
    bulk_g2s(smem_addr(cb), cq + (int64_t)sb * np + row0, cnt16, bar_c);
  };
  __syncthreads();
  if (tid == 0 && it.valid()) load_cells(it);
  for (; it.valid();) {
    const int role = it.role();
    const int64_t tile = it.tile();
    if (role != cur_role) {
      // park the sums of the role we leave in its slot, fetch the sums of the one we enter
      if (!EMIT) {
        if (cur_role >= 0) {
          double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
          for (int e = 0; e < kCellAcc; ++e) dst[e * kCellThreads + tid] = acc[e];
        }
        const double* src = my_part + (int64_t)(role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
        for (int e = 0; e < kCellAcc; ++e) acc[e] = src[e * kCellThreads + tid];
      }
      if (tid == 0) R = roles[role];
      cur_role = role;
    }
    const int64_t row0 = tile * TR;
    const int cnt = (int)min((int64_t)TR, n - row0);
    if (EMIT) {
      if (tid == 0) s_careful = 0;
    } else if (tid == 0) {
      // t of both terms and omega: in flight while the rows are being sorted.  The workspace columns are
      // 16-byte aligned and padded to 16 rows: the last tile is rounded up
      const uint32_t cnt16 = (uint32_t)((cnt + 15) & ~15);
      const DCellRole* gr = roles + role;
      const double* ga = tq + (int64_t)gr->sa * np + row0;
      const double* gb = gr->diag ? omz + row0 : tq + (int64_t)gr->sb * np + row0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar_b, 3u * cnt16 * 8u);
      bulk_g2s(smem_addr(ta), ga, cnt16 * 8u, bar_b);
      bulk_g2s(smem_addr(tb), gb, cnt16 * 8u, bar_b);
      bulk_g2s(smem_addr(wv), om + row0, cnt16 * 8u, bar_b);
      s_careful = 0;
    }
    {
      uint4* rec = reinterpret_cast<uint4*>(wh32 + tid * W);
      rec[0] = make_uint4(0u, 0u, 0u, 0u);
      if (W == 8) rec[1] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const int NC = R.ncell, diag = R.diag, ncb = R.ncb, key0 = R.key0;
    mbar_wait(bar_c, phase);
    // ---- counting sort of the tile's rows by cell, stable per warp.  Counters are private to a
    //      (warp, cell): 16 bits each, two per word, bumped with one shared-memory atomic on the word
    //      (a warp's conflicting lanes are resolved by the hardware in a fixed order, and no other warp
    //      touches the word: the sorted order, hence the fp64 sums, are reproducible run to run).
    //      Rows of other roles of the same pair get key 0xffff.  A thread takes four consecutive rows.
    const uint32_t wq = (uint32_t)warp >> 1, winc = (warp & 1) ? 0x10000u : 1u;
    const uint32_t hmul = R.hmul;
    int careful = 0;
    for (int r4 = 4 * tid; r4 < cnt; r4 += 4 * kCellThreads) {
      const uint32_t a4 = *reinterpret_cast<const uint32_t*>(ca + r4);
      const uint32_t b4 = *reinterpret_cast<const uint32_t*>(cb + r4);
      careful |= (int)((a4 | (diag ? 0u : b4)) & 0x80808080u);
      uint32_t kk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ci = (a4 >> (8 * q)) & 0x7f;
        const int cj = diag ? (int)(((uint32_t)(r4 + q) * hmul) >> 16) : (int)((b4 >> (8 * q)) & 0x7f);
        const int k = ci * ncb + cj - key0;
        const bool mine = ((unsigned)k < (unsigned)NC) && (r4 + q < cnt);
        kk[q] = mine ? (uint32_t)k : 0xffffu;
        if (mine) atomicAdd(&wh32[k * W + ((wq + ((uint32_t)k >> SH)) & (W - 1))], winc);
      }
      *reinterpret_cast<uint32_t*>(ca + r4) = kk[0] | (kk[1] << 16);
      *reinterpret_cast<uint32_t*>(cb + r4) = kk[2] | (kk[3] << 16);
    }
    if (careful) s_careful = 1;
    __syncthreads();
    // ---- thread k owns cell k: its rows per warp, exclusive scan over the cells, start offsets back
    uint32_t cw[W];
    {
      const uint4* rec = reinterpret_cast<const uint4*>(wh32 + tid * W);
      const uint4 r0 = rec[0];
      cw[0] = r0.x; cw[1] = r0.y; cw[2] = r0.z; cw[3] = r0.w;
      if (W == 8) {
        const uint4 r1 = rec[1];
        cw[4 % W] = r1.x; cw[5 % W] = r1.y; cw[6 % W] = r1.z; cw[7 % W] = r1.w;
      }
    }
    int c = 0;
#pragma unroll
    for (int q = 0; q < W; ++q) c += (int)((cw[q] & 0xffffu) + (cw[q] >> 16));
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int t = lane < kCellWarps ? s_wtot[lane] : 0;
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += v;
      }
      s_wtot[lane] = ti - t;
    }
    __syncthreads();
    const int start = incl - c + s_wtot[warp];
    {
      // the counters become the start offsets of every (cell, warp) run, in storage order
      uint32_t run = (uint32_t)start;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        const uint32_t lo = cw[q] & 0xffffu, hi = cw[q] >> 16;
        cw[q] = run | ((run + lo) << 16);
        run += lo + hi;
      }
      uint4* rec = reinterpret_cast<uint4*>(wh32 + tid * W);
      rec[0] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      if (W == 8) rec[1] = make_uint4(cw[4 % W], cw[5 % W], cw[6 % W], cw[7 % W]);
    }
    __syncthreads();
    for (int r4 = 4 * tid; r4 < cnt; r4 += 4 * kCellThreads) {
      const uint32_t k01 = *reinterpret_cast<const uint32_t*>(ca + r4), k23 = *reinterpret_cast<const uint32_t*>(cb + r4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t k = (((q & 2) ? k23 : k01) >> (16 * (q & 1))) & 0xffffu;
        if (k != 0xffffu) {
          const uint32_t old = atomicAdd(&wh32[k * W + ((wq + (k >> SH)) & (W - 1))], winc);
          perm[(warp & 1) ? (old >> 16) : (old & 0xffffu)] = (uint16_t)(r4 + q);
        }
      }
    }
    __syncthreads();
    // ---- the interval bytes of the next unit go where the keys were, while this unit's rows are summed
    it.next();
    if (tid == 0 && it.valid()) load_cells(it);
    if (EMIT) {
      // the order of this (role, tile): sorted row indices, start of every cell, total, "careful" flag
      const int64_t ntl = (n + TR - 1) / TR;
      uint32_t* pg = reinterpret_cast<uint32_t*>(perm_g + ((int64_t)role * ntl + tile) * TR);
      const uint32_t* ps = reinterpret_cast<const uint32_t*>(perm);
      for (int i = tid; i < (cnt + 1) / 2; i += kCellThreads) pg[i] = ps[i];
      uint16_t* sg = starts_g + ((int64_t)role * ntl + tile) * (kCellThreads + 8);
      sg[tid] = (uint16_t)start;
      if (tid == kCellThreads - 1) {
        sg[kCellThreads] = (uint16_t)(start + c);
        sg[kCellThreads + 1] = (uint16_t)(s_careful ? 1 : 0);
      }
      phase ^= 1;
      __syncthreads();
      continue;
    }
    mbar_wait(bar_b, phase);
    phase ^= 1;
    // ---- the rows of my cell
    if (!s_careful) {
      if (!diag) cell_rows_pair<false>(&R, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<false>(&R, perm + start, c, ta, tb, wv, acc);
    } else {
      if (!diag) cell_rows_pair<true>(&R, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<true>(&R, perm + start, c, ta, tb, wv, acc);
    }
    __syncthreads();
  }
  if (!EMIT && cur_role >= 0) {
    double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
    for (int e = 0; e < kCellAcc; ++e) dst[e * kCellThreads + tid] = acc[e];
  }
}

// The cells kernel of a prepared shard (pgb_gram_prepare): the rows of every (role, tile) are already in cell
// order, so a tile is one set of bulk copies (t_i, t_j, omega, the sorted row indices, the cell starts) and the
// sums.  Same scheduling, partial slots and reproducible order as gram_cells_kernel.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
gram_cells_sorted_kernel(int64_t n, int64_t np, const double* __restrict__ om, const double* __restrict__ omz,
                         const double* __restrict__ tq, const DCellRole* __restrict__ roles, int n_roles, int TR,
                         int band_tiles, int slots_per_cta, double* __restrict__ part,
                         const uint16_t* __restrict__ perm_g, const uint16_t* __restrict__ starts_g) {
  constexpr int kCellThreads = THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ DCellRole R;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x;
  double* ta = reinterpret_cast<double*>(smem_raw);
  double* tb = ta + TR;
  double* wv = tb + TR;
  uint16_t* perm = reinterpret_cast<uint16_t*>(wv + TR);
  uint16_t* starts = perm + TR;  // [THREADS + 8]
  const uint32_t bar = smem_addr(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  const int64_t ntl = (n + TR - 1) / TR;
  const int role_lo = (int)((int64_t)n_roles * blockIdx.x / gridDim.x);
  double* my_part = part + (int64_t)blockIdx.x * slots_per_cta * (kCellAcc * kCellThreads);
  int cur_role = -1;
  uint32_t phase = 0;
  double acc[kCellAcc];
#pragma unroll
  for (int e = 0; e < kCellAcc; ++e) acc[e] = 0.0;
  UnitIter it;
  it.init(n, TR, band_tiles, n_roles);
  __syncthreads();
  for (; it.valid(); it.next()) {
    const int role = it.role();
    const int64_t tile = it.tile();
    const int64_t row0 = tile * TR;
    const int cnt = (int)min((int64_t)TR, n - row0);
    if (tid == 0) {
      const uint32_t cnt16 = (uint32_t)((cnt + 15) & ~15);
      const DCellRole* gr = roles + role;
      const double* ga = tq + (int64_t)gr->sa * np + row0;
      const double* gb = gr->diag ? omz + row0 : tq + (int64_t)gr->sb * np + row0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous tile's generic reads vs the async writes
      mbar_expect_tx(bar, 3u * cnt16 * 8u + cnt16 * 2u + (uint32_t)(kCellThreads + 8) * 2u);
      bulk_g2s(smem_addr(ta), ga, cnt16 * 8u, bar);
      bulk_g2s(smem_addr(tb), gb, cnt16 * 8u, bar);
      bulk_g2s(smem_addr(wv), om + row0, cnt16 * 8u, bar);
      bulk_g2s(smem_addr(perm), perm_g + ((int64_t)role * ntl + tile) * TR, cnt16 * 2u, bar);
      bulk_g2s(smem_addr(starts), starts_g + ((int64_t)role * ntl + tile) * (kCellThreads + 8), (uint32_t)(kCellThreads + 8) * 2u, bar);
    }
    if (role != cur_role) {
      if (cur_role >= 0) {
        double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
        for (int e = 0; e < kCellAcc; ++e) dst[e * kCellThreads + tid] = acc[e];
      }
      const double* src = my_part + (int64_t)(role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
      for (int e = 0; e < kCellAcc; ++e) acc[e] = src[e * kCellThreads + tid];
      if (tid == 0) R = roles[role];
      cur_role = role;
      __syncthreads();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    const int start = starts[tid], c = (int)starts[tid + 1] - start;
    const int diag = R.diag;
    if (!starts[kCellThreads + 1]) {
      if (!diag) cell_rows_pair<false>(&R, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<false>(&R, perm + start, c, ta, tb, wv, acc);
    } else {
      if (!diag) cell_rows_pair<true>(&R, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<true>(&R, perm + start, c, ta, tb, wv, acc);
    }
    __syncthreads();
  }
  if (cur_role >= 0) {
    double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * kCellThreads);
#pragma unroll
    for (int e = 0; e < kCellAcc; ++e) dst[e * kCellThreads + tid] = acc[e];
  }
}

// ----------------------------------------------------------------------------------------
// The sum kernel of a prepared shard as a RING: one CTA per SM runs G independent groups; a group = NW warps
// (thread = cell, as above) and a ring of kRingStages tile buffers with a full / empty
// mbarrier pair each.  Lane 0 of the group's first warp issues the five bulk copies of the NEXT unit (t_i, t_j, omega, the
// sorted row indices, the cell starts) as soon as every consumer warp has released the stage; a consumer warp waits
// for `full`, sums the rows of its 32 cells and arrives on `empty` -- no CTA-wide barrier anywhere in the loop.
// Why (ncu source page of the barrier version, profiles/r2_ncu_c3_sorted_320x2_source.txt): 25 % of its stall samples
// sat in the wait for a tile's copies (issued and awaited by the same CTA, hidden only by the second CTA of the SM)
// and 9 % behind the barrier that ends a tile, where every warp waits for the fullest of the CTA's 512 cells; here a
// warp waits for the fullest of its own 32 cells only and the copies of the next unit are in flight while this one
// is summed.
// ----------------------------------------------------------------------------------------

template <int THREADS, int G>
__global__ void __launch_bounds__(G * THREADS, 1)
gram_cells_ring_kernel(int64_t n, int64_t np, const double* __restrict__ om, const double* __restrict__ omz,
                       const double* __restrict__ tq, const DCellRole* __restrict__ roles, int n_roles, int TR,
                       int band_tiles, int slots_per_cta, double* __restrict__ part,
                       const uint16_t* __restrict__ perm_g, const uint16_t* __restrict__ starts_g) {
  constexpr int NW = THREADS / 32, S = kRingStages;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[G][2 * S];
  const int g = threadIdx.x / THREADS, t = threadIdx.x - g * THREADS, lane = t & 31;
  const size_t stage_bytes = (((size_t)TR * 26 + (size_t)(THREADS + 8) * 2) + 15) & ~(size_t)15;
  unsigned char* gbase = smem_raw + (size_t)g * S * stage_bytes;
  if (t == 0) {
    for (int q = 0; q < S; ++q) {
      mbar_init(smem_addr(&s_bar[g][q]), 1);           // full: the issuing lane's arrive + the bytes of the copies
      mbar_init(smem_addr(&s_bar[g][S + q]), NW);      // empty: one arrive per warp
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int64_t ntl = (n + TR - 1) / TR;
  const int vcta = (int)blockIdx.x * G + g, nvcta = (int)gridDim.x * G;
  UnitIter it;
  it.init(n, TR, band_tiles, n_roles, vcta, nvcta);
  // the copies of unit j are issued by lane 0 of warp 0 when it STARTS unit j - 1 (unit 0: before the loop), as soon
  // as every warp has released the stage of unit j - 2: they have a whole unit of summing to arrive
  const bool issuer = (t == 0);
  UnitIter ahead = it;
  auto issue = [&](const UnitIter& u, int j) {
    const int q = j % S;
    if (j >= S) mbar_wait(smem_addr(&s_bar[g][S + q]), (uint32_t)((j / S) - 1) & 1u);
    const int role = u.role();
    const int64_t tile = u.tile(), row0 = tile * TR;
    const uint32_t cnt16 = (uint32_t)((min((int64_t)TR, n - row0) + 15) & ~(int64_t)15);
    const DCellRole* gr = roles + role;
    double* ta = reinterpret_cast<double*>(gbase + (size_t)q * stage_bytes);
    const double* ga = tq + (int64_t)gr->sa * np + row0;
    const double* gb = gr->diag ? omz + row0 : tq + (int64_t)gr->sb * np + row0;
    const uint32_t bar = smem_addr(&s_bar[g][q]);
    mbar_expect_tx(bar, 3u * cnt16 * 8u + cnt16 * 2u + (uint32_t)(THREADS + 8) * 2u);
    bulk_g2s(smem_addr(ta), ga, cnt16 * 8u, bar);
    bulk_g2s(smem_addr(ta + TR), gb, cnt16 * 8u, bar);
    bulk_g2s(smem_addr(ta + 2 * TR), om + row0, cnt16 * 8u, bar);
    bulk_g2s(smem_addr(ta + 3 * TR), perm_g + ((int64_t)role * ntl + tile) * TR, cnt16 * 2u, bar);
    bulk_g2s(smem_addr(reinterpret_cast<uint16_t*>(ta + 3 * TR) + TR), starts_g + ((int64_t)role * ntl + tile) * (THREADS + 8),
             (uint32_t)(THREADS + 8) * 2u, bar);
  };
  if (issuer && ahead.valid()) issue(ahead, 0);
  if (ahead.valid()) ahead.next();
  // thread t owns cell key0 + t of the role it serves
  const int role_lo = (int)((int64_t)n_roles * vcta / nvcta);
  double* my_part = part + (int64_t)vcta * slots_per_cta * (kCellAcc * THREADS);
  int cur_role = -1, diag = 0;
  const DCellRole* Rg = roles;
  double acc[kCellAcc];
#pragma unroll
  for (int e = 0; e < kCellAcc; ++e) acc[e] = 0.0;
  for (int i = 0; it.valid(); it.next(), ++i) {
    if (ahead.valid()) {
      if (issuer) issue(ahead, i + 1);
      ahead.next();
    }
    const int q = i % S, role = it.role();
    if (role != cur_role) {
      if (cur_role >= 0) {
        double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * THREADS);
#pragma unroll
        for (int e = 0; e < kCellAcc; ++e) dst[e * THREADS + t] = acc[e];
      }
      const double* src = my_part + (int64_t)(role - role_lo) * (kCellAcc * THREADS);
#pragma unroll
      for (int e = 0; e < kCellAcc; ++e) acc[e] = src[e * THREADS + t];
      Rg = roles + role;
      diag = Rg->diag;
      cur_role = role;
    }
    mbar_wait(smem_addr(&s_bar[g][q]), (uint32_t)(i / S) & 1u);
    const unsigned char* sb = gbase + (size_t)q * stage_bytes;
    const double* ta = reinterpret_cast<const double*>(sb);
    const double* tb = ta + TR;
    const double* wv = tb + TR;
    const uint16_t* perm = reinterpret_cast<const uint16_t*>(wv + TR);
    const uint16_t* starts = perm + TR;
    const int start = starts[t], c = (int)starts[t + 1] - start;
    if (!starts[THREADS + 1]) {
      if (!diag) cell_rows_pair<false>(Rg, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<false>(Rg, perm + start, c, ta, tb, wv, acc);
    } else {
      if (!diag) cell_rows_pair<true>(Rg, perm + start, c, ta, tb, wv, acc);
      else cell_rows_diag<true>(Rg, perm + start, c, ta, tb, wv, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_addr(&s_bar[g][S + q]));
  }
  if (cur_role >= 0) {
    double* dst = my_part + (int64_t)(cur_role - role_lo) * (kCellAcc * THREADS);
#pragma unroll
    for (int e = 0; e < kCellAcc; ++e) dst[e * THREADS + t] = acc[e];
  }
}

// the prepass of a prepared shard: lp from the stored cell records (no basis interval search, no record
// writes), then the IRLS step; writes omega and omega z and the iteration scalars
__global__ void __launch_bounds__(kPreThreads, 3)
gram_prepass_planned_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam, int mode, int64_t n,
                            const double* __restrict__ y, const float* __restrict__ w, const double* __restrict__ beta,
                            int64_t np, const double* __restrict__ tq, const uint8_t* __restrict__ cq,
                            double* __restrict__ om, double* __restrict__ omz, double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + (PGB_GS_COUNT + 2) * 32);
  PreSpl* s_spl = reinterpret_cast<PreSpl*>(s_terms + n_terms);
  __shared__ int s_nspl, s_icpt;
  for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();
  if (tid == 0) {
    int ns = 0, ic = -1;
    for (int q = 0; q < n_terms; ++q) {
      const DTerm& T = s_terms[q];
      if (T.kind == PGB_TERM_INTERCEPT) { ic = T.coef_offset; continue; }
      const DMarg& mg = T.marg[0];
      s_spl[ns++] = PreSpl{mg.lo, mg.inv_span, (double)mg.nint, mg.feature, mg.nint, T.coef_offset, q};
    }
    s_nspl = ns;
    s_icpt = ic;
  }
  __syncthreads();
  const int n_spl = s_nspl;
  const double icpt_beta = s_icpt >= 0 ? s_beta[s_icpt] : 0.0;
  double sc[PGB_GS_COUNT + 2];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT + 2; ++s) sc[s] = 0.0;
  constexpr int NX = 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yi = __ldg(y + i);
    double o, oz;
    if (mode == PGB_MODE_IRLS) {
      const double winv = weight_inv(w, i);
      double lp = icpt_beta;
      for (int base = 0; base < n_spl; base += NX) {
        double tv[NX];
        int cv[NX];
#pragma unroll
        for (int q = 0; q < NX; ++q)
          if (base + q < n_spl) {
            tv[q] = __ldg(tq + (int64_t)(base + q) * np + i);
            cv[q] = __ldg(cq + (int64_t)(base + q) * np + i);
          }
#pragma unroll
        for (int q = 0; q < NX; ++q) {
          if (base + q >= n_spl) break;
          const PreSpl& S = s_spl[base + q];
          double v[4];
          if (!(cv[q] & 0x80)) {
            cubic_bspline(tv[q] + 0.5, v[0], v[1], v[2], v[3]);
          } else {
            const double u = tv[q] < 0.0 ? tv[q] + 1.0 : tv[q] - 1.0;
            extrap_basis(s_terms[S.term].marg[0], u, v);
          }
          const double* bt = s_beta + S.coef_offset + (cv[q] & 0x7f);
          lp = fma(v[0], bt[0], lp);
          lp = fma(v[1], bt[1], lp);
          lp = fma(v[2], bt[2], lp);
          lp = fma(v[3], bt[3], lp);
        }
      }
      const IrlsRow ir = irls_row(fam, lp, yi, winv);
      o = ir.omega;
      oz = ir.omega * ir.z;
      if (ir.keep) {
        sc[PGB_GS_NUSED] += 1.0;
        sc[PGB_GS_DEV] += dist_dev(fam, yi, ir.mu);
        sc[PGB_GS_NCORRECT] += (yi == ((ir.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      o = 1.0;
      oz = init_response(fam, yi);
      sc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(oz)) sc[PGB_GS_NONFINITE] += 1.0;
    }
    sc[PGB_GS_NROWS] += 1.0;
    sc[PGB_GS_COUNT] += o;
    sc[PGB_GS_COUNT + 1] += oz;
    om[i] = o;
    omz[i] = oz;
  }
  block_reduce_store<PGB_GS_COUNT + 2>(sc, s_red, scal_part + (int64_t)blockIdx.x * (PGB_GS_COUNT + 2));
}

// ----------------------------------------------------------------------------------------
// models with te() / f() / l() terms: the prepass of every term kind, the sort of the generic orders and the
// generic sum kernel
// ----------------------------------------------------------------------------------------
// One thread per row: lp over all terms, the IRLS step, and the cell record of every factor column: cubic
// marginals (s(), both marginals of te()) as in gram_prepass_kernel, f(): the coefficient index of the row's
// level (0xff: none), l(): the feature value.  te() rows follow TensorTerm.build_columns (terms.py:1454-1477,
// utils.py:899-948: index a * K_2 + b), f() rows FactorTerm.build_columns (terms.py:1077-1096), l() rows
// LinearTerm.build_columns (terms.py:693-708).
__global__ void __launch_bounds__(kPreThreads, 3)
gram_prepass_mixed_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam, int mode,
                          const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                          const float* __restrict__ w, const double* __restrict__ beta, int64_t np, double* __restrict__ om,
                          double* __restrict__ omz, double* __restrict__ tq, uint8_t* __restrict__ cq,
                          double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + (PGB_GS_COUNT + 2) * 32);
  for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();
  double sc[PGB_GS_COUNT + 2];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT + 2; ++s) sc[s] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yi = __ldg(y + i);
    const double winv = weight_inv(w, i);
    const double lp = prepass_row_lp(
        s_terms, n_terms, s_beta, [&](int f) { return __ldg(X + (int64_t)f * ld + i); },
        [&](int c, double t) { tq[(int64_t)c * np + i] = t; }, [&](int c, int b) { cq[(int64_t)c * np + i] = (uint8_t)b; });
    double o, oz;
    if (mode == PGB_MODE_IRLS) {
      const IrlsRow ir = irls_row(fam, lp, yi, winv);
      o = ir.omega;
      oz = ir.omega * ir.z;
      if (ir.keep) {
        sc[PGB_GS_NUSED] += 1.0;
        sc[PGB_GS_DEV] += dist_dev(fam, yi, ir.mu);
        sc[PGB_GS_NCORRECT] += (yi == ((ir.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      o = 1.0;
      oz = init_response(fam, yi);
      sc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(oz)) sc[PGB_GS_NONFINITE] += 1.0;
    }
    sc[PGB_GS_NROWS] += 1.0;
    sc[PGB_GS_COUNT] += o;       // G[intercept, intercept]
    sc[PGB_GS_COUNT + 1] += oz;  // r[intercept]
    om[i] = o;
    omz[i] = oz;
  }
  block_reduce_store<PGB_GS_COUNT + 2>(sc, s_red, scal_part + (int64_t)blockIdx.x * (PGB_GS_COUNT + 2));
}

// The order of a generic key: per (order, tile) the tile's rows sorted by cell (ties by row index, so the order and
// with it every fp64 sum is reproducible) and the start of every cell.  Counting sort in shared memory, kGenChunk
// cells at a time; rows without an entry (level byte 0xff) are left out.
static const int kGenThreads = 256;
static const int kGenChunk = 2048;
__global__ void __launch_bounds__(kGenThreads)
gram_cells_order_kernel(int64_t n, int64_t np, const uint8_t* __restrict__ cq, const DGenOrder* __restrict__ orders,
                        int n_orders, int TR, int64_t ntl, uint16_t* __restrict__ perm_g, uint16_t* __restrict__ starts_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* keys = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* perm = keys + TR;
  int* cnt = reinterpret_cast<int*>(perm + TR);
  int* cur = cnt + kGenChunk;
  __shared__ DGenOrder O;
  __shared__ int s_wtot[kGenThreads / 32], s_base, s_careful;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PER = kGenChunk / kGenThreads;
  for (int64_t u = blockIdx.x; u < ntl * n_orders; u += gridDim.x) {
    const int64_t tile = u / n_orders;
    const int o = (int)(u - tile * n_orders);
    __syncthreads();
    if (tid == 0) { O = orders[o]; s_base = 0; s_careful = 0; }
    __syncthreads();
    const int64_t row0 = tile * TR;
    const int cnt_rows = (int)min((int64_t)TR, n - row0);
    uint32_t outside = 0;  // a row outside the edge knots of a cubic key column: the sum kernel takes its careful variant
    for (int r = tid; r < cnt_rows; r += kGenThreads)
      keys[r] = (uint16_t)gen_key(O, r, [&](int c) {
        const uint32_t b = __ldg(cq + (int64_t)c * np + row0 + r);
        outside |= b;
        return b;
      });
    bool any_out = false;
    for (int d = 0; d < O.nkey; ++d) any_out |= O.key_kind[d] == PGB_FK_CUBIC;
    if (any_out && (outside & 0x80u)) s_careful = 1;
    uint16_t* sg = starts_g + O.starts_off * ntl + tile * O.spitch;
    for (int k0 = 0; k0 < O.ncells; k0 += kGenChunk) {
      for (int j = tid; j < kGenChunk; j += kGenThreads) cnt[j] = 0;
      __syncthreads();
      for (int r = tid; r < cnt_rows; r += kGenThreads) {
        const unsigned k = (unsigned)keys[r] - (unsigned)k0;
        if (keys[r] != 0xffffu && k < (unsigned)kGenChunk) atomicAdd(&cnt[k], 1);
      }
      __syncthreads();
      // exclusive scan of the chunk's counters: thread t owns counters PER t .. PER t + PER - 1
      int loc[PER], tot = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) { loc[j] = tot; tot += cnt[tid * PER + j]; }
      int incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      int wbase = 0;
      for (int q = 0; q < warp; ++q) wbase += s_wtot[q];
      const int base = s_base + wbase + incl - tot;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int k = k0 + tid * PER + j;
        cur[tid * PER + j] = base + loc[j];
        cnt[tid * PER + j] = base + loc[j];
        if (k < O.ncells) sg[k] = (uint16_t)(base + loc[j]);
      }
      __syncthreads();
      if (tid == kGenThreads - 1) s_base = base + tot;
      for (int r = tid; r < cnt_rows; r += kGenThreads) {
        const unsigned k = (unsigned)keys[r] - (unsigned)k0;
        if (keys[r] != 0xffffu && k < (unsigned)kGenChunk) perm[atomicAdd(&cur[k], 1)] = (uint16_t)r;
      }
      __syncthreads();
      // rows of a cell in ascending order (the atomics above place them in an arbitrary one)
      for (int j = tid; j < kGenChunk && k0 + j < O.ncells; j += kGenThreads) {
        const int a = cnt[j], b = cur[j];
        for (int i = a + 1; i < b; ++i) {
          const uint16_t v = perm[i];
          int q = i - 1;
          while (q >= a && perm[q] > v) { perm[q + 1] = perm[q]; --q; }
          perm[q + 1] = v;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      sg[O.ncells] = (uint16_t)s_base;
      sg[O.ncells + 1] = (uint16_t)s_careful;
    }
    const int placed = s_base;
    uint16_t* pg = perm_g + ((int64_t)o * ntl + tile) * TR;
    for (int i = tid; i < placed; i += kGenThreads) pg[i] = perm[i];
  }
}

// ---- weighted units of the generic kernel.  A role's tile costs about (rows per cell + a constant) dependent
// gathers, which differs by an order of magnitude between a 4-d te() x te() pair and the intercept column of a te():
// role r counts w_r virtual units per tile, the virtual axis of a band (nb tiles x sum of the weights) is cut into
// equal shares, and a (role, tile) belongs to the CTA whose share holds its first virtual unit.
struct WUnitIter {
  const int32_t* pre;  // prefix sums of the role weights (n_roles + 1 entries, shared memory)
  int64_t ntiles, nbands, band, nb, t0, v_lo, v_hi;
  int64_t t;
  int band_tiles, n_roles, W, r;
  bool ok;
  __device__ int find_role(int64_t x) const {  // largest r with pre[r] * nb <= x
    int lo = 0, hi = n_roles - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((int64_t)pre[mid] * nb <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
  }
  __device__ int64_t vstart() const { return (int64_t)pre[r] * nb + t * (pre[r + 1] - pre[r]); }
  __device__ void set_band() {
    t0 = band * band_tiles;
    nb = min((int64_t)band_tiles, ntiles - t0);
    const int64_t V = nb * W;
    v_lo = V * blockIdx.x / gridDim.x;
    v_hi = V * (blockIdx.x + 1) / gridDim.x;
    ok = false;
    if (v_hi <= v_lo) return;
    if (!(band & 1)) {  // forwards: first unit that starts at or after v_lo
      r = find_role(v_lo);
      const int w = pre[r + 1] - pre[r];
      t = (v_lo - (int64_t)pre[r] * nb + w - 1) / w;
      if (t >= nb) { ++r; t = 0; }
      ok = r < n_roles && vstart() < v_hi;
    } else {  // backwards: last unit that starts before v_hi
      r = find_role(v_hi - 1);
      const int w = pre[r + 1] - pre[r];
      t = min(nb - 1, (v_hi - 1 - (int64_t)pre[r] * nb) / w);
      ok = vstart() >= v_lo;
    }
  }
  __device__ void skip_empty() {
    while (band < nbands && !ok) {
      ++band;
      if (band < nbands) set_band();
    }
  }
  __device__ void init(int64_t n, int TR, int bt, int nr, const int32_t* prefix) {
    pre = prefix;
    band_tiles = bt;
    n_roles = nr;
    W = prefix[nr];
    ntiles = (n + TR - 1) / TR;
    nbands = (ntiles + bt - 1) / bt;
    band = 0;
    set_band();
    skip_empty();
  }
  __device__ bool valid() const { return band < nbands; }
  __device__ int role() const { return r; }
  __device__ int64_t tile() const { return t0 + t; }
  __device__ void next() {
    if (!(band & 1)) {
      if (++t >= nb) { ++r; t = 0; }
      ok = r < n_roles && vstart() < v_hi;
    } else {
      if (--t < 0) { --r; t = nb - 1; }
      ok = r >= 0 && vstart() >= v_lo;
    }
    skip_empty();
  }
};

template <bool CAREFUL>
__device__ __forceinline__ void gen_pw(int nint, double s, double (&p)[4]) {
  if (CAREFUL) {
    record_powers(nint, s, p);
  } else {
    p[0] = 1.0;
    p[1] = s;
    p[2] = s * s;
    p[3] = p[2] * s;
  }
}
// s^e for a thread-constant exponent e
template <bool CAREFUL>
__device__ __forceinline__ double gen_pow(int nint, double s, int e) {
  if (CAREFUL) {
    double p[4];
    record_powers(nint, s, p);
    return sel4(p, e);
  }
  const double x = (e & 1) ? s : 1.0, y = (e & 2) ? s * s : 1.0;
  return x * y;
}

// inner block of a row: acc[4 p + q] += (mult a_p) b_q (NIN = 2), acc[q] += mult b_q (NIN = 1), acc[0] += mult
template <int NIN, bool CAREFUL>
__device__ __forceinline__ void gen_inner(double mult, const double (&a)[4], const double (&b)[4], double* acc) {
  if (NIN == 0) {
    acc[0] += mult;
  } else if (NIN == 1) {
    if (CAREFUL) acc[0] = fma(mult, b[0], acc[0]); else acc[0] += mult;
    acc[1] = fma(mult, b[1], acc[1]); acc[2] = fma(mult, b[2], acc[2]); acc[3] = fma(mult, b[3], acc[3]);
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const double u = (p == 0 && !CAREFUL) ? mult : mult * a[p];
      if (CAREFUL) acc[4 * p] = fma(u, b[0], acc[4 * p]); else acc[4 * p] += u;
      acc[4 * p + 1] = fma(u, b[1], acc[4 * p + 1]);
      acc[4 * p + 2] = fma(u, b[2], acc[4 * p + 2]);
      acc[4 * p + 3] = fma(u, b[3], acc[4 * p + 3]);
    }
  }
}

// One row of a generic pair: NOUT outer cubic factors, NIN inner ones, NVAL value factors.  With outer factors the
// thread covers all four exponents of the first one (sums acc[16 j + ..] for exponent j) and the exponent e1 of its
// slice of the second.  v[] = the row's record coordinates in the order (values, outer 0, outer 1, inner a, inner b).
template <int NOUT, int NIN, int NVAL, bool CAREFUL>
__device__ __forceinline__ void gen_row(double mult, const double* v, const int* nint, int e1, double (&acc)[kGenAcc]) {
#pragma unroll
  for (int q = 0; q < NVAL; ++q) mult *= v[q];
  if (NOUT >= 2) mult *= gen_pow<CAREFUL>(nint[1], v[NVAL + 1], e1);
  double a[4] = {1.0, 0.0, 0.0, 0.0}, b[4] = {1.0, 0.0, 0.0, 0.0};
  if (NIN == 2) gen_pw<CAREFUL>(nint[2], v[NVAL + NOUT], a);
  if (NIN >= 1) gen_pw<CAREFUL>(nint[3], v[NVAL + NOUT + NIN - 1], b);
  if (NOUT == 0) {
    gen_inner<NIN, CAREFUL>(mult, a, b, acc);
  } else {
    double o[4];
    gen_pw<CAREFUL>(nint[0], v[NVAL], o);
#pragma unroll
    for (int j = 0; j < 4; ++j) gen_inner<NIN, CAREFUL>((j == 0 && !CAREFUL) ? mult : mult * o[j], a, b, acc + 16 * j);
  }
}

// the rows s0 .. s1 - 1 of a cell in the stored order, two rows in flight: the gathers of a pair of rows are issued
// together, the row indices of the next pair are fetched one step ahead
template <int NOUT, int NIN, int NVAL, bool CAREFUL>
__device__ __forceinline__ void gen_rows(const uint16_t* __restrict__ pg, int s0, int s1, const double* __restrict__ wcol,
                                         const double* const* cols, const int* nint, int e1, double (&acc)[kGenAcc]) {
  constexpr int NC = NVAL + NOUT + NIN;
  int r0 = pg[s0], r1 = (s0 + 1 < s1) ? pg[s0 + 1] : 0;
  for (int s = s0; s < s1; s += 2) {
    const bool two = s + 1 < s1;
    const int ra = r0, rb = r1;
    if (s + 2 < s1) r0 = pg[s + 2];
    if (s + 3 < s1) r1 = pg[s + 3];
    double va[NC > 0 ? NC : 1], vb[NC > 0 ? NC : 1];
    const double wa = __ldg(wcol + ra);
    const double wb = two ? __ldg(wcol + rb) : 0.0;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      va[q] = __ldg(cols[q] + ra);
      vb[q] = two ? __ldg(cols[q] + rb) : 0.0;
    }
    gen_row<NOUT, NIN, NVAL, CAREFUL>(wa, va, nint, e1, acc);
    if (two) gen_row<NOUT, NIN, NVAL, CAREFUL>(wb, vb, nint, e1, acc);
  }
}

// The generic sum kernel: thread (cell, slice) of a role walks the rows of its cell in the stored order and reads
// their record coordinates straight from the (L2-resident) band -- a role touches only the few rows of its own
// cells, so nothing is staged in shared memory; the lanes of a cell's slices read the same addresses.  Same
// partial slots and reproducible order as the spline kernels.
__global__ void __launch_bounds__(kGenSumThreads, kGenOcc)
gram_cells_generic_kernel(int64_t n, int64_t np, const double* __restrict__ om, const double* __restrict__ omz,
                          const double* __restrict__ tq, const DGenRole* __restrict__ roles,
                          const DGenPair* __restrict__ pairs, int n_roles, const int32_t* __restrict__ role_prefix,
                          const int32_t* __restrict__ cta_role_lo, int TR, int band_tiles, int slots_per_cta,
                          double* __restrict__ part, const uint16_t* __restrict__ perm_g,
                          const uint16_t* __restrict__ starts_g, int64_t ntl) {
  constexpr int THREADS = kGenSumThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* s_pre = reinterpret_cast<int32_t*>(smem_raw);
  __shared__ DGenPair P;
  __shared__ DGenRole R;
  const int tid = threadIdx.x;
  for (int i = tid; i <= n_roles; i += THREADS) s_pre[i] = role_prefix[i];
  __syncthreads();
  const int role_lo = cta_role_lo[blockIdx.x];
  double* my_part = part + (int64_t)blockIdx.x * slots_per_cta * (kGenAcc * THREADS);
  int cur_role = -1;
  double acc[kGenAcc];
#pragma unroll
  for (int e = 0; e < kGenAcc; ++e) acc[e] = 0.0;
  // per role: this thread's cell and slice, the columns it gathers
  int k = 0, e1 = 0, shape = 0;
  int nint[4] = {1, 1, 1, 1};
  const double* cbase[6] = {tq, tq, tq, tq, tq, tq};
  WUnitIter it;
  it.init(n, TR, band_tiles, n_roles, s_pre);
  for (; it.valid(); it.next()) {
    const int role = it.role();
    const int64_t tile = it.tile();
    if (role != cur_role) {
      if (cur_role >= 0) {
        double* dst = my_part + (int64_t)(cur_role - role_lo) * (kGenAcc * THREADS);
#pragma unroll
        for (int e = 0; e < kGenAcc; ++e) dst[e * THREADS + tid] = acc[e];
      }
      const double* src = my_part + (int64_t)(role - role_lo) * (kGenAcc * THREADS);
#pragma unroll
      for (int e = 0; e < kGenAcc; ++e) acc[e] = src[e * THREADS + tid];
      __syncthreads();
      if (tid == 0) {
        R = roles[role];
        P = pairs[R.pair];
      }
      __syncthreads();
      cur_role = role;
      const int nslt = P.nslt;
      k = tid / nslt;
      e1 = tid - k * nslt;
      const int nout = (P.out0 >= 0) + (P.out1 >= 0), nin = (P.in_b >= 0) + (P.in_a >= 0);
      shape = nout * 9 + nin * 3 + P.nval;
      int q = 0;
      for (int j = 0; j < P.nval; ++j) cbase[q++] = tq + (int64_t)P.val_col[j] * np;
      if (P.out0 >= 0) { nint[0] = P.f_n[P.out0]; cbase[q++] = tq + (int64_t)P.f_col[P.out0] * np; }
      if (P.out1 >= 0) { nint[1] = P.f_n[P.out1]; cbase[q++] = tq + (int64_t)P.f_col[P.out1] * np; }
      if (P.in_a >= 0) { nint[2] = P.f_n[P.in_a]; cbase[q++] = tq + (int64_t)P.f_col[P.in_a] * np; }
      if (P.in_b >= 0) { nint[3] = P.f_n[P.in_b]; cbase[q++] = tq + (int64_t)P.f_col[P.in_b] * np; }
      if (k >= R.ncl) k = -1;
    }
    if (k < 0) continue;
    const uint16_t* sgt = starts_g + P.starts_off * ntl + tile * P.spitch;
    const int s0 = sgt[R.key0 + k], s1 = sgt[R.key0 + k + 1];
    if (s0 >= s1) continue;
    const bool careful = sgt[P.ncells + 1] != 0;
    const int64_t row0 = tile * TR;
    const uint16_t* pg = perm_g + ((int64_t)P.order * ntl + tile) * TR;
    const double* wcol = (P.wz ? omz : om) + row0;
    const double* cols[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) cols[q] = cbase[q] + row0;
#define PGB_GEN_CASE(NOUT, NIN, NVAL)                                                                             \
  case (NOUT) * 9 + (NIN) * 3 + (NVAL):                                                                           \
    if (!careful) gen_rows<NOUT, NIN, NVAL, false>(pg, s0, s1, wcol, cols, nint, e1, acc);                        \
    else gen_rows<NOUT, NIN, NVAL, true>(pg, s0, s1, wcol, cols, nint, e1, acc);                                  \
    break;
    switch (shape) {
      PGB_GEN_CASE(1, 2, 0)  // s() x te()
      PGB_GEN_CASE(2, 2, 0)  // te() x te(), the diagonal block of a te()
      PGB_GEN_CASE(0, 2, 0)  // te() x intercept / rhs / f()
      PGB_GEN_CASE(0, 2, 1)  // te() x l()
      PGB_GEN_CASE(0, 1, 0)  // s() x f()
      PGB_GEN_CASE(0, 1, 1)  // s() x l()
      PGB_GEN_CASE(0, 0, 0)  // f() x f() / intercept / rhs
      PGB_GEN_CASE(0, 0, 1)  // l() x f() / intercept / rhs
      PGB_GEN_CASE(0, 0, 2)  // l() x l()
      PGB_GEN_CASE(0, 2, 2)  // s(by=) x s(by=) / te(s, l): a cubic and a value factor on either side
      PGB_GEN_CASE(1, 2, 1)  // s(by=) / te(s, l) x te(s, s)
      PGB_GEN_CASE(0, 1, 2)  // s(by=) / te(s, l) x l()
      default: break;        // no other shape is planned (pgb_plan_cells: cells_shape_supported)
    }
#undef PGB_GEN_CASE
  }
  if (cur_role >= 0) {
    double* dst = my_part + (int64_t)(cur_role - role_lo) * (kGenAcc * THREADS);
#pragma unroll
    for (int e = 0; e < kGenAcc; ++e) dst[e * THREADS + tid] = acc[e];
  }
}

// generic sums[pair][product][cell] = sum over the CTAs that served the cell's role, fixed order; product =
// sum index of the thread + n_acc * slice
__global__ void cells_generic_reduce_kernel(const double* __restrict__ part, const DGenRole* __restrict__ roles,
                                            const DGenPair* __restrict__ pairs, const int32_t* __restrict__ cta_role_lo,
                                            int slots_per_cta, double* __restrict__ sums) {
  const int role = blockIdx.x, tid = threadIdx.x, e = blockIdx.y;
  const DGenRole R = roles[role];
  const DGenPair& P = pairs[R.pair];
  const int nslt = P.nslt, nthr = P.n_acc * P.ne;  // sums per thread
  const int k = tid / nslt, sg = tid - k * nslt;
  if (k >= R.ncl || e >= nthr) return;
  // thread sum e = inner index + 16 j (j: exponent of the first outer factor; 16 = n_acc whenever ne == 4)
  double v = 0.0;
  for (int c = R.c_lo; c <= R.c_hi; ++c) {
    const int slot = role - cta_role_lo[c];
    if (slot < 0 || slot >= slots_per_cta) continue;
    v += part[((int64_t)c * slots_per_cta + slot) * (kGenAcc * kGenSumThreads) + e * kGenSumThreads + tid];
  }
  sums[P.sum_off + (int64_t)(e + nthr * sg) * P.ncells + R.key0 + k] = v;
}

// change of basis of the generic sums along moment axis `axis`: one thread per line of four sums
__global__ void cells_generic_transform_kernel(const DGenPair* __restrict__ pairs, int axis, double* __restrict__ sums) {
  const DGenPair& P = pairs[blockIdx.y];
  if (axis >= P.n_ax) return;
  const int64_t lines = (int64_t)P.ncells * (P.nprod / 4);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= lines) return;
  const int64_t cell = idx % P.ncells, l = idx / P.ncells;
  const int64_t sh = (int64_t)1 << (2 * axis);
  const int64_t e0 = (l / sh) * (4 * sh) + (l % sh);
  moment_line(sums + P.sum_off + e0 * P.ncells + cell, sh * P.ncells);
}

// ----------------------------------------------------------------------------------------
// kernel 3: sums[pair][product][cell] = sum over the CTAs that served the cell's role, fixed order
// ----------------------------------------------------------------------------------------
__global__ void cells_slot_reduce_kernel(const double* __restrict__ part, const DCellRole* __restrict__ roles,
                                         const DCellPair* __restrict__ pairs, int n_roles, int grid, int slots_per_cta,
                                         int threads, double* __restrict__ sums) {
  const int role = blockIdx.x, k = threadIdx.x, e = blockIdx.y;
  const DCellRole R = roles[role];
  if (k >= R.ncell) return;
  const DCellPair P = pairs[R.pair];
  const int64_t nr = n_roles;
  const int c_lo = R.c_lo, c_hi = R.c_hi;
  {
    double v = 0.0;
    for (int c = c_lo; c <= c_hi; ++c) {
      const int lo = (int)(nr * c / grid);
      const int hi = (int)((nr * (c + 1) + grid - 1) / grid) - 1;
      if (role < lo || role > hi || role - lo >= slots_per_cta) continue;
      v += part[((int64_t)c * slots_per_cta + (role - lo)) * (kCellAcc * threads) + e * threads + k];
    }
    sums[P.sum_off + (int64_t)e * P.ncells + R.key0 + k] = v;
  }
}

// moment sums of a cell -> the entries of G the cell touches, in place: B = M S M^T (pair), the packed
// upper triangle of M S M^T with S symmetric (diagonal block), M x for the intercept column and the rhs
__global__ void cells_transform_kernel(const DCellPair* __restrict__ pairs, int n_pairs, double* __restrict__ sums) {
  const DCellPair P = pairs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.ncells) return;
  spline_transform_cell(P.tj < 0, sums + P.sum_off + k, P.ncells);
}

// scalars of the prepass: the public ones into out, sum(omega) and sum(omega z) for the assembly.
// One warp per scalar: lane l sums the partials l, l + 32, ..., then a shuffle tree (fixed order).
__global__ void cells_scalar_reduce_kernel(const double* __restrict__ scal_part, int count, double* __restrict__ out_scal,
                                           double* __restrict__ extra, int accumulate) {
  const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int c = lane; c < count; c += 32) v += scal_part[(int64_t)c * (PGB_GS_COUNT + 2) + s];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane != 0) return;
  if (s < PGB_GS_COUNT) out_scal[s] = accumulate ? out_scal[s] + v : v;
  else extra[s - PGB_GS_COUNT] = v;
}

// ----------------------------------------------------------------------------------------
// kernel 4: [G | r] from the cell sums.  One thread per entry (row <= col) of the upper triangle.
// ----------------------------------------------------------------------------------------
__global__ void cells_assemble_kernel(const double* __restrict__ sums, const double* __restrict__ extra, CellsView V,
                                      double* __restrict__ out, int accumulate) {
  const int m = V.m;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * (m + 1)) return;
  const int r = (int)(idx / (m + 1)), c = (int)(idx - (int64_t)r * (m + 1));
  if (c < r) return;
  double v = 0.0;
  cells_entry_visit(V, r, c, [&](int64_t off, int cnt) {
    if (off < 0) {
      v += extra[off == -1 ? 0 : 1];
    } else {
      const double* q = sums + off;
      for (int h = 0; h < cnt; ++h) v += q[h];
    }
  });
  if (c == m) {
    double* dst = out + (int64_t)m * m + r;
    *dst = accumulate ? *dst + v : v;
  } else {
    double* a = out + (int64_t)r * m + c;
    const double nv = accumulate ? *a + v : v;
    *a = nv;
    out[(int64_t)c * m + r] = nv;
  }
}

struct CellWs {
  double *part, *gpart, *sums, *scal_part, *extra, *om, *omz, *tq;
  uint8_t* cq;
  uint16_t *perm, *starts, *gperm, *gstarts;
  int64_t np, ntl, gntl;
};
static bool cells_ws_layout(const pgb_model* M, int64_t n, void* ws, int64_t ws_bytes, bool with_order, CellWs& L) {
  const pgb_model::Cells& C = M->cells;
  L.np = rows_pad(n);
  L.ntl = (n + C.tile_rows - 1) / C.tile_rows;
  L.gntl = C.mixed ? (n + C.gtile_rows - 1) / C.gtile_rows : 0;
  L.part = (double*)ws;
  L.gpart = L.part + (int64_t)C.grid * C.slots_per_cta * C.slot_doubles;
  L.sums = L.gpart + (int64_t)C.ggrid * C.gslots_per_cta * C.gslot_doubles;
  L.scal_part = L.sums + C.sum_doubles + C.gsum_doubles;
  L.extra = L.scal_part + (int64_t)kSMs * 8 * (PGB_GS_COUNT + 2);
  L.om = (double*)(((uintptr_t)(L.extra + 8) + 127) & ~(uintptr_t)127);
  L.omz = L.om + L.np;
  L.tq = L.omz + L.np;
  L.cq = (uint8_t*)(L.tq + (int64_t)C.ncol * L.np);
  L.perm = (uint16_t*)(((uintptr_t)(L.cq + (int64_t)C.ncol * L.np) + 127) & ~(uintptr_t)127);
  L.starts = L.perm + (int64_t)C.roles.size() * L.ntl * C.tile_rows;
  L.gperm = (uint16_t*)(((uintptr_t)(L.starts + (int64_t)C.roles.size() * L.ntl * (C.threads + 8)) + 127) & ~(uintptr_t)127);
  L.gstarts = L.gperm + (int64_t)C.gorders.size() * L.gntl * C.gtile_rows;
  const uint8_t* end = (with_order || C.mixed) ? (const uint8_t*)(L.gstarts + C.gstarts_u16 * L.gntl) : (const uint8_t*)L.perm;
  return (int64_t)(end - (const uint8_t*)ws) <= ws_bytes;
}

static size_t cells_pre_smem(const pgb_model* M) {
  return ((size_t)((M->m + 1) & ~1) * 8 + (PGB_GS_COUNT + 2) * 32 * 8 + (size_t)M->n_terms * (sizeof(DTerm) + sizeof(PreSpl)) + 127) &
         ~(size_t)127;
}
static int cells_band_tiles(const pgb_model* M) {
  const pgb_model::Cells& C = M->cells;
  const int64_t row_bytes = 9 * (int64_t)C.ncol + 16;
  int band_tiles = (int)std::max<int64_t>(1, kBandBytes / (row_bytes * C.tile_rows));
  if (M->dbg.cells_band_tiles > 0) band_tiles = M->dbg.cells_band_tiles;
  return band_tiles;
}

#define PGB_CELLS_DISPATCH(KERNEL_CALL)                 \
  switch (C.threads) {                                  \
    case 256: KERNEL_CALL(256, 2); break;               \
    case 320: KERNEL_CALL(320, 2); break;               \
    case 384: KERNEL_CALL(384, 1); break;               \
    default: KERNEL_CALL(512, 1); break;                \
  }

static int cells_set_smem_attr(const void* fn, int bytes) {
  // per device (a process may drive several GPUs): the attribute is cheap to set, so no caching
  PGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return PGB_OK;
}

// the sort kernels: cell order of every (spline role, tile) and of every (generic order, tile), from the cell records
static int cells_emit_order(const pgb_model* M, int64_t n, const CellWs& L, cudaStream_t st) {
  const pgb_model::Cells& C = M->cells;
  const int n_roles = (int)C.roles.size();
  const int TR = C.tile_rows, band_tiles = cells_band_tiles(M);
  if (n_roles > 0) {
#define PGB_CELLS_EMIT(TH, OCC)                                                                                          \
  do {                                                                                                                   \
    int rc_ = cells_set_smem_attr((const void*)gram_cells_kernel<TH, OCC, true>,                                         \
                                  (int)(OCC == 2 ? kSmemMax / 2 - 2048 : kSmemMax - 1024));                               \
    if (rc_ != PGB_OK) return rc_;                                                                                       \
    gram_cells_kernel<TH, OCC, true><<<C.grid, TH, C.smem, st>>>(n, L.np, L.om, L.omz, L.tq, L.cq, C.d_roles, n_roles, TR, \
                                                                 band_tiles, C.slots_per_cta, L.part, L.perm, L.starts); \
  } while (0)
    PGB_CELLS_DISPATCH(PGB_CELLS_EMIT)
#undef PGB_CELLS_EMIT
    PGB_CUDA(cudaGetLastError());
    pgb_count_launch(1);
  }
  if (!C.gorders.empty()) {
    const int smem = 4 * C.gtile_rows + 8 * kGenChunk;
    int rc = cells_set_smem_attr((const void*)gram_cells_order_kernel, smem);
    if (rc != PGB_OK) return rc;
    const int64_t units = L.gntl * (int64_t)C.gorders.size();
    gram_cells_order_kernel<<<(unsigned)std::min<int64_t>(units, (int64_t)kSMs * 4), kGenThreads, smem, st>>>(
        n, L.np, L.cq, C.d_gorders, (int)C.gorders.size(), C.gtile_rows, L.gntl, L.gperm, L.gstarts);
    PGB_CUDA(cudaGetLastError());
    pgb_count_launch(1);
  }
  return PGB_OK;
}

static int cells_launch_prepass(const pgb_model* M, int32_t mode, bool planned, const double* X, int64_t n, int64_t ld,
                                const double* y, const float* w, const double* beta, const CellWs& L, int& grid1,
                                cudaStream_t st) {
  const pgb_model::Cells& C = M->cells;
  grid1 = (int)std::min<int64_t>((n + kPreThreads - 1) / kPreThreads, (int64_t)kSMs * 8);
  const size_t pre_smem = cells_pre_smem(M);
  if (planned && mode == PGB_MODE_IRLS) {
    // a prepared shard needs only omega and omega z of every row: the O(k) pass of pgb_core.cu (TMA-staged columns, one
    // cubic per knot interval from a conflict-free table) does that at twice the rate of the record kernels below
    const int rc = pgb_pass_irls(M, X, n, ld, y, w, beta, L.om, L.omz, L.scal_part, &grid1, st);
    if (rc != PGB_EUNSUPPORTED) return rc;
    grid1 = (int)std::min<int64_t>((n + kPreThreads - 1) / kPreThreads, (int64_t)kSMs * 8);
  }
  if (C.mixed)
    gram_prepass_mixed_kernel<<<grid1, kPreThreads, pre_smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, mode, X, n, ld, y, w, beta,
                                                                    L.np, L.om, L.omz, L.tq, L.cq, L.scal_part);
  else if (planned)
    gram_prepass_planned_kernel<<<grid1, kPreThreads, pre_smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, mode, n, y, w, beta, L.np,
                                                                      L.tq, L.cq, L.om, L.omz, L.scal_part);
  else
    gram_prepass_kernel<<<grid1, kPreThreads, pre_smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, mode, X, n, ld, y, w, beta, L.np,
                                                              L.om, L.omz, L.tq, L.cq, L.scal_part);
  PGB_CUDA(cudaGetLastError());
  pgb_count_launch(1);
  return PGB_OK;
}

// pgb_gram_prepare: cell records of every row and the cell order of every (role, tile)
int pgb_cells_prepare(const pgb_model* M, const double* X, int64_t n, int64_t ld, void* ws, int64_t ws_bytes, cudaStream_t st) {
  CellWs L;
  if (!cells_ws_layout(M, n, ws, ws_bytes, true, L)) {
    pgb_set_error("pgb_gram_prepare: workspace too small");
    return PGB_EINVAL;
  }
  // records only: the response of this launch is column 0 of X (any n finite doubles will do), its omega is discarded
  int grid1 = 0;
  int rc = cells_launch_prepass(M, PGB_MODE_INIT, false, X, n, ld, X, nullptr, nullptr, L, grid1, st);
  if (rc != PGB_OK) return rc;
  return cells_emit_order(M, n, L, st);
}

// pgb_gram_refine: see gram_cells_refine_kernel
int pgb_cells_refine(const pgb_model* M, int64_t n, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const pgb_model::Cells& C = M->cells;
  CellWs L;
  if (!cells_ws_layout(M, n, ws, ws_bytes, true, L)) {
    pgb_set_error("pgb_gram_refine: workspace too small");
    return PGB_EINVAL;
  }
  const int n_roles = (int)C.roles.size();
  if (n_roles == 0) return PGB_OK;
  const int TR = C.tile_rows;
  const unsigned grid = (unsigned)std::min<int64_t>(L.ntl * n_roles, (int64_t)kSMs * 8);
  switch (C.threads) {
    case 256: gram_cells_refine_kernel<256><<<grid, 256, (size_t)TR * 2 + 64, st>>>(n, TR, n_roles, L.perm, L.starts); break;
    case 320: gram_cells_refine_kernel<320><<<grid, 320, (size_t)TR * 2 + 64, st>>>(n, TR, n_roles, L.perm, L.starts); break;
    case 384: gram_cells_refine_kernel<384><<<grid, 384, (size_t)TR * 2 + 64, st>>>(n, TR, n_roles, L.perm, L.starts); break;
    default: gram_cells_refine_kernel<512><<<grid, 512, (size_t)TR * 2 + 64, st>>>(n, TR, n_roles, L.perm, L.starts); break;
  }
  PGB_CUDA(cudaGetLastError());
  pgb_count_launch(1);
  return PGB_OK;
}

int pgb_launch_gram_cells(const pgb_model* M, int32_t mode, const double* X, int64_t n, int64_t ld, const double* y,
                          const float* w, const double* beta, double* out, void* ws, int64_t ws_bytes, int accumulate,
                          int planned, cudaStream_t st) {
  const pgb_model::Cells& C = M->cells;
  pgb_model* Mm = const_cast<pgb_model*>(M);
  const int n_roles = (int)C.roles.size(), n_groles = (int)C.groles.size();
  CellWs L;
  if (!cells_ws_layout(M, n, ws, ws_bytes, planned != 0, L)) {
    pgb_set_error("pgb_gram_pass: workspace too small");
    return PGB_EINVAL;
  }
  double *part = L.part, *sums = L.sums, *scal_part = L.scal_part, *extra = L.extra, *om = L.om, *omz = L.omz, *tq = L.tq;
  uint8_t* cq = L.cq;
  const int64_t np = L.np;
  pgb_model::Timing& tm = Mm->timing;
  if (tm.enabled) {
    while ((int)tm.chunk_ev.size() < 5) {
      cudaEvent_t e;
      PGB_CUDA(cudaEventCreate(&e));
      tm.chunk_ev.push_back(e);
    }
    PGB_CUDA(cudaEventRecord(tm.chunk_ev[0], st));
    PGB_CUDA(cudaEventRecord(tm.chunk_ev[2], st));
  }
  int grid1 = 0;
  int rc = cells_launch_prepass(M, mode, planned != 0, X, n, ld, y, w, beta, L, grid1, st);
  if (rc != PGB_OK) return rc;
  // a model with generic pairs that was not prepared sorts its rows now (rows streamed through once)
  if (C.mixed && !planned) {
    rc = cells_emit_order(M, n, L, st);
    if (rc != PGB_OK) return rc;
  }
  PGB_CUDA(cudaMemsetAsync(part, 0, sizeof(double) * ((int64_t)C.grid * C.slots_per_cta * C.slot_doubles + (int64_t)C.ggrid * C.gslots_per_cta * C.gslot_doubles), st));
  if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.chunk_ev[3], st));
  const int TR = C.tile_rows, band_tiles = cells_band_tiles(M);
  const bool sorted = planned || C.mixed;
#define PGB_CELLS_LAUNCH(TH, OCC)                                                                                        \
  do {                                                                                                                   \
    if (n_roles > 0 && sorted && C.ring_groups > 0) {                                                                    \
      constexpr int G_ = (TH <= 384) ? 2 : 1;                                                                            \
      const size_t smem_ = (size_t)G_ * kRingStages * ring_stage_bytes(TR, TH);                                          \
      rc = cells_set_smem_attr((const void*)gram_cells_ring_kernel<TH, G_>, (int)(kSmemMax - 1024));                     \
      if (rc != PGB_OK) return rc;                                                                                       \
      gram_cells_ring_kernel<TH, G_><<<kSMs, G_ * TH, smem_, st>>>(n, np, om, omz, tq, C.d_roles, n_roles, TR,           \
                                                                   band_tiles, C.slots_per_cta, part, L.perm, L.starts); \
    } else if (n_roles > 0 && sorted) {                                                                                  \
      rc = cells_set_smem_attr((const void*)gram_cells_sorted_kernel<TH, 2>, (int)(kSmemMax / 2 - 2048));                \
      if (rc != PGB_OK) return rc;                                                                                       \
      gram_cells_sorted_kernel<TH, 2><<<C.grid, TH, (size_t)TR * 26 + (TH + 8) * 2 + 64, st>>>(                          \
          n, np, om, omz, tq, C.d_roles, n_roles, TR, band_tiles, C.slots_per_cta, part, L.perm, L.starts);              \
    } else if (n_roles > 0) {                                                                                            \
      rc = cells_set_smem_attr((const void*)gram_cells_kernel<TH, OCC, false>,                                           \
                               (int)(OCC == 2 ? kSmemMax / 2 - 2048 : kSmemMax - 1024));                                  \
      if (rc != PGB_OK) return rc;                                                                                       \
      gram_cells_kernel<TH, OCC, false><<<C.grid, TH, C.smem, st>>>(n, np, om, omz, tq, cq, C.d_roles, n_roles, TR,      \
                                                                    band_tiles, C.slots_per_cta, part, nullptr, nullptr); \
    }                                                                                                                    \
  } while (0)
  PGB_CELLS_DISPATCH(PGB_CELLS_LAUNCH)
#undef PGB_CELLS_LAUNCH
  if (n_groles > 0)
    gram_cells_generic_kernel<<<C.ggrid, kGenSumThreads, (size_t)(n_groles + 1) * 4, st>>>(
        n, np, om, omz, tq, C.d_groles, C.d_gpairs, n_groles, C.d_grole_prefix, C.d_gcta_role_lo, C.gtile_rows,
        std::max(1, (int)((int64_t)band_tiles * TR / C.gtile_rows)), C.gslots_per_cta, L.gpart, L.gperm, L.gstarts, L.gntl);
  PGB_CUDA(cudaGetLastError());
  if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.chunk_ev[4], st));
  int launches = 3;
  if (n_roles > 0) {
    cells_slot_reduce_kernel<<<dim3(n_roles, kCellAcc), C.threads, 0, st>>>(part, C.d_roles, C.d_pairs, n_roles, C.grid,
                                                                          C.slots_per_cta, C.threads, sums);
    int max_cells = 1;
    for (const DCellPair& P : C.pairs) max_cells = std::max(max_cells, P.ncells);
    cells_transform_kernel<<<dim3((max_cells + 127) / 128, (unsigned)C.pairs.size()), 128, 0, st>>>(C.d_pairs, (int)C.pairs.size(), sums);
    launches += 3;
  }
  if (n_groles > 0) {
    cells_generic_reduce_kernel<<<dim3(n_groles, kGenAcc), kGenSumThreads, 0, st>>>(L.gpart, C.d_groles, C.d_gpairs, C.d_gcta_role_lo,
                                                                                   C.gslots_per_cta, sums);
    int64_t max_lines = 1;
    int max_ax = 0;
    for (const DGenPair& P : C.gpairs) {
      max_lines = std::max(max_lines, (int64_t)P.ncells * (P.nprod / 4));
      max_ax = std::max(max_ax, P.n_ax);
    }
    for (int ax = 0; ax < max_ax; ++ax)
      cells_generic_transform_kernel<<<dim3((unsigned)((max_lines + 127) / 128), (unsigned)C.gpairs.size()), 128, 0, st>>>(C.d_gpairs, ax, sums);
    launches += 2 + max_ax;
  }
  cells_scalar_reduce_kernel<<<1, 32 * (PGB_GS_COUNT + 2), 0, st>>>(scal_part, grid1, out + (int64_t)M->m * M->m + M->m, extra,
                                                                     accumulate);
  const int64_t entries = (int64_t)M->m * (M->m + 1);
  const CellsView V{C.d_coef_term, C.d_pair_of, C.d_term_class, C.d_gpair_of, C.d_pairs, C.d_gpairs, M->n_terms, M->m, C.icpt_coef};
  cells_assemble_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, st>>>(sums, extra, V, out, accumulate);
  PGB_CUDA(cudaGetLastError());
  pgb_count_launch(launches);
  if (tm.enabled) {
    PGB_CUDA(cudaEventRecord(tm.chunk_ev[1], st));
    tm.pending = true;
    tm.n_chunks = 1;
  }
  return PGB_OK;
}
