// pgb_gram.cu -- the fused Gram pass of one PIRLS iteration.
//
// Replaces pygam.py:764-783 of the reference (lp -> mu -> W -> mask -> pseudo data -> W.B and
// the n-sized part of its QR) and the data part of _initial_estimate (pygam.py:696-710):
//     G = X^T W^2 X,   r = X^T W^2 z,   plus the per-iteration log scalars,
// with the model-matrix rows evaluated on the fly (never materialised).
//
// Work decomposition (pgb_internal.h): the record slots of a row are grouped into quads, the
// coefficients into groups; G[group gi, group gj] is a BLOCK kept in shared memory by one role
// (CTA), with a row pitch = 4 (mod 8) doubles.  An ITEM = (row quad, column quad): the 16 lanes
// (a, b) of a half-warp add value[a] * value[b] at (index[a], index[b]) -- 4 consecutive rows x
// 4 consecutive columns = 16 distinct 8-byte banks, no conflicts.  A warp owns its items (two
// per step, one per half-warp) for the whole kernel, so the fp64 read-modify-write needs no
// atomics.  Phase A stages a tile of rows as 16-byte (value, index) slot records (one thread per
// (row, term): basis, link, weight); phase B walks the tile.  The pass is bound by the
// shared-memory read-modify-write rate (16 B per FMA at 128 B/clk/SM), see DESIGN.md.  Each CTA
// writes its accumulators to a private partial; gram_reduce_kernel sums the partials in a fixed
// order (bit-reproducible) and mirrors the upper triangle.
#include <cuda_runtime.h>
#include <stdio.h>

#include <algorithm>

#include "pgb_internal.h"
#include "pgb_kernels.cuh"

static const int kMaxWarps = 15;      // consumer (accumulating) warps per CTA of the accumulation kernel; + 1 TMA warp
static const int kNBuf = 3;           // record buffers of the accumulation kernel (TMA ring)
static const int kSlotBytes = 10;     // record slot: double value + uint16 index (two arrays)
static const int kRecThreads = 256;   // CTA of the records kernel
static const int64_t kChunkRows = 1 << 22;  // about this many rows' records are materialised at a time

// ----------------------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------------------
static const int kMinTile = 16;
// row pitch of the slot-major tile records ([slot][pitch]): four spare rows (the accumulation works
// on groups of four rows and prefetches one group ahead) and, with tiles that are multiples of 8 rows,
// pitch = 4 (mod 8): the four slots of a quad start in different banks, every slot row is 16-byte
// aligned (TMA bulk copies, 128-bit value loads) and four 16-bit indices are one aligned 64-bit word
static int tile_pitch(int tile) { return tile + 4; }

static int pitch_for(int width) {
  int p = std::max(width, 4);
  while ((p & 7) != 4) ++p;
  return p;
}

// doubles of a stored block: rows = the varying group, pitch = the shared (column-indexing) group
static int64_t block_doubles(const pgb_model* M, const DBlock& b) {
  const DGroup &gv = M->groups[b.tr ? b.gj : b.gi], &gs = M->groups[b.tr ? b.gi : b.gj];
  return ((int64_t)gv.width * gs.pitch + 1) & ~(int64_t)1;
}

static size_t role_fixed_smem(const pgb_model* M) {
  return (size_t)M->n_terms * sizeof(DTerm) + (size_t)M->m * 8 + PGB_GS_COUNT * 32 * 8 + 512;
}

int pgb_plan_gram(pgb_model* M) {
  const int m = M->m, nt = M->n_terms;
  M->pass_ctas = kSMs * 8;  // needed by the O(k) passes whether or not this plan succeeds
  std::vector<DTerm>& T = M->dterms;
  // ---- groups and quads; gcoef[g][idx] = coefficient (m = rhs) behind index idx of group g
  M->groups.clear();
  std::vector<std::vector<int>> gcoef;
  std::vector<int> quad_real;  // real (non-padding) slots of every quad
  std::vector<int> quad_group;
  int quads = 0;
  for (int j = 0; j < nt; ++j) {
    if (T[j].nnz < 2) continue;
    DGroup g{};
    g.width = T[j].n_coefs;
    g.n_quads = (T[j].nnz + 3) / 4;
    g.quad0 = quads;
    for (int q = 0; q < g.n_quads; ++q) {
      quad_real.push_back(std::min(4, T[j].nnz - 4 * q));
      quad_group.push_back((int)M->groups.size());
    }
    quads += g.n_quads;
    T[j].group = (int)M->groups.size();
    T[j].idx_base = 0;
    T[j].gslot0 = 4 * g.quad0;
    std::vector<int> cf(g.width);
    for (int i = 0; i < g.width; ++i) cf[i] = T[j].coef_offset + i;
    gcoef.push_back(cf);
    M->groups.push_back(g);
  }
  // single-slot terms (intercept, l(), f(), ...) and the virtual rhs column share misc groups
  int open = -1, used = 0;
  auto misc_slot = [&]() {
    if (open < 0 || used == 4) {
      DGroup g{};
      g.n_quads = 1;
      g.quad0 = quads++;
      quad_real.push_back(0);
      quad_group.push_back((int)M->groups.size());
      open = (int)M->groups.size();
      used = 0;
      gcoef.push_back(std::vector<int>());
      M->groups.push_back(g);
    }
    ++quad_real[M->groups[open].quad0];
    return used++;
  };
  for (int j = 0; j < nt; ++j) {
    if (T[j].nnz >= 2) continue;
    const int s = misc_slot();
    DGroup& g = M->groups[open];
    T[j].group = open;
    T[j].idx_base = g.width;
    T[j].gslot0 = 4 * g.quad0 + s;
    for (int i = 0; i < T[j].n_coefs; ++i) gcoef[open].push_back(T[j].coef_offset + i);
    g.width += T[j].n_coefs;
  }
  const int rhs_s = misc_slot();
  const int rhs_group = open, rhs_idx = M->groups[open].width, rhs_quad = M->groups[open].quad0;
  gcoef[open].push_back(m);
  M->groups[open].width += 1;
  const int ng = (int)M->groups.size();
  for (DGroup& g : M->groups) g.pitch = pitch_for(g.width);

  // ---- blocks.  The groups are cut into PANELS of consecutive groups about sqrt(budget)
  //      coefficients wide and the blocks are ordered by (panel of gi, panel of gj): a role, a run of
  //      consecutive blocks, then covers about one panel pair and needs the record slots of two
  //      panels only -- less shared memory for records, larger tiles (a 2-d decomposition of G).
  const size_t fixed = role_fixed_smem(M);
  const size_t rec_row_all = (size_t)(quads * 4) * kSlotBytes;  // bytes per tile row when a role needs every quad
  if ((kMinTile + 4) * rec_row_all * 2 + 32768 > kSmemMax) {
    pgb_set_error("model too wide for the Gram pass: k=%d non-zeros per row", M->k);
    return PGB_EUNSUPPORTED;
  }
  int best_tile = 0, n_roles = 0;
  double best_budget = 0.0;
  int tile = 0;
  size_t smem = 0;
  int max_warps = 1;
  const int est_slots = std::min(4 * quads, std::max(32, 2 * quads));
  double budget_d = (double)(kSmemMax - (size_t)kNBuf * 36 * est_slots * kSlotBytes - 8192) / 8.0;
  for (int attempt = 0; attempt < 14; ++attempt) {
  const bool final_pass = (attempt == 13);
  if (final_pass) {
    if (best_tile == 0) {
      pgb_set_error("Gram pass does not fit in shared memory");
      return PGB_EUNSUPPORTED;
    }
    budget_d = best_budget;
  }
  const int64_t acc_budget = (int64_t)budget_d;
  std::vector<int> panel(ng, 0);
  {
    const double B = std::sqrt((double)acc_budget);
    int p = 0;
    double wsum = 0.0;
    for (int g = 0; g < ng; ++g) {
      const double wg = (double)M->groups[g].pitch;
      if (wsum > 0.0 && wsum + wg > B) { ++p; wsum = 0.0; }
      panel[g] = p;
      wsum += wg;
    }
  }
  M->blocks.clear();
  int64_t total_acc = 0;
  bool block_too_big = false;
  for (int gi = 0; gi < ng; ++gi)
    for (int gj = gi; gj < ng; ++gj) {
      DBlock b{};
      b.gi = gi;
      b.gj = gj;
      const int qi = M->groups[gi].n_quads, qj = M->groups[gj].n_quads;
      b.n_items = (gi == gj) ? qi * (qi + 1) / 2 : qi * qj;
      // the quad whose operand is shared by the items of a half-step indexes the columns of the
      // stored block: the side with fewer quads (the row group => the block is stored transposed)
      b.tr = (gi != gj && b.n_items > 1 && qi <= qj) ? 1 : 0;
      const int64_t sz = block_doubles(M, b);
      if (sz > acc_budget) block_too_big = true;
      total_acc += sz;
      M->blocks.push_back(b);
    }
  if (block_too_big) {
    if (attempt == 0) {
      pgb_set_error("a Gram block does not fit in shared memory (m=%d)", m);
      return PGB_EUNSUPPORTED;
    }
    attempt = 12;  // smaller budgets cannot work either: take the best so far
    continue;
  }
  std::stable_sort(M->blocks.begin(), M->blocks.end(), [&](const DBlock& x, const DBlock& y) {
    if (panel[x.gi] != panel[y.gi]) return panel[x.gi] < panel[y.gi];
    return panel[x.gj] < panel[y.gj];
  });
  // roles of equal accumulator size: walk the blocks in order and close a role when it reaches
  // its share; more roles if one of them overflows the shared-memory budget
  std::vector<int64_t> role_acc;
  for (int want = (int)((total_acc + acc_budget - 1) / acc_budget);; ++want) {
    const int64_t target = (total_acc + want - 1) / want;
    role_acc.clear();
    int role = 0;
    int64_t acc = 0, cum = 0;
    bool ok = true;
    for (DBlock& b : M->blocks) {
      const int64_t sz = block_doubles(M, b);
      const int at = (int)std::min<int64_t>(want - 1, (cum + sz / 2) / target);
      if (acc > 0 && (at > role || acc + sz > acc_budget)) {
        role_acc.push_back(acc);
        ++role;
        acc = 0;
      }
      b.role = role;
      b.off = acc;
      acc += sz;
      cum += sz;
      if (acc > acc_budget) ok = false;
    }
    role_acc.push_back(acc);
    if (ok) break;
  }
  n_roles = (int)role_acc.size();

  // ---- per role: record layout, terms to evaluate, items -> steps -> warps
  M->roles.clear();
  M->rterms.clear();
  M->steps.clear();
  M->warps.clear();
  M->runs.clear();
  M->n_slots_all = 4 * quads;
  max_warps = 1;
  for (int r = 0; r < n_roles; ++r) {
    DRole R{};
    R.acc_size = role_acc[r];
    std::vector<char> gneed(ng, 0);
    for (const DBlock& b : M->blocks)
      if (b.role == r) gneed[b.gi] = gneed[b.gj] = 1;
    std::vector<int> lquad(quads, -1);  // global quad -> role-local quad
    int nlq = 0;
    bool all = true;
    for (int g = 0; g < ng; ++g) {
      if (!gneed[g]) { all = false; continue; }
      for (int q = 0; q < M->groups[g].n_quads; ++q) lquad[M->groups[g].quad0 + q] = nlq++;
    }
    R.n_slots = 4 * nlq;
    R.needs_all_terms = all ? 1 : 0;
    // contiguous runs of global quads -> one TMA bulk copy each (values, indices) per tile
    R.run0 = (int)M->runs.size();
    for (int q = 0; q < quads; ++q) {
      if (lquad[q] < 0) continue;
      if (!M->runs.empty() && (int)M->runs.size() > R.run0 && M->runs.back().gslot0 + M->runs.back().n_slots == 4 * q)
        M->runs.back().n_slots += 4;
      else
        M->runs.push_back(DRun{4 * q, 4 * lquad[q], 4, 0});
    }
    R.n_runs = (int)M->runs.size() - R.run0;
    R.rhs_slot = gneed[rhs_group] ? 4 * lquad[rhs_quad] + rhs_s : -1;
    // Items -> half-steps -> chains.  A block with several items (tensor terms, splines of order
    // > 3) can reach the same accumulator element through different quad pairs on different data
    // rows, so all its half-steps form one chain that stays in one half-warp column of one warp
    // (the items of one half-step work on the same data row: distinct elements).  Single-item
    // blocks are bundled by column quad: the column operand is loaded once per data row.
    struct Item { int rq, cq; const DBlock* b; };
    // bit 4 * (varying slot) + (shared slot): the lane accumulates; bit 16: same group on both sides
    auto item_mask = [&](int rq, int cq, bool same_group, bool shared_is_row) {
      uint32_t mk = same_group ? (1u << 16) : 0u;
      for (int a = 0; a < quad_real[rq]; ++a) {
        if (rq == rhs_quad && a == rhs_s) continue;  // the rhs is a column only
        for (int c = 0; c < quad_real[cq]; ++c)
          if (rq != cq || a <= c) mk |= 1u << (shared_is_row ? 4 * c + a : 4 * a + c);
      }
      return mk;
    };
    auto make_half = [&](const std::vector<Item>& its, bool shared_is_row) {
      DHalf h{};
      h.shared_slot = (uint32_t)(4 * lquad[shared_is_row ? its[0].rq : its[0].cq]);
      for (size_t k = 0; k < its.size(); ++k) {
        const Item& it = its[k];
        h.v[k].slot = (uint32_t)(4 * lquad[shared_is_row ? it.cq : it.rq]);
        h.v[k].off8 = (uint32_t)(it.b->off * 8);
        h.v[k].pitch8 = (uint32_t)(M->groups[shared_is_row ? it.b->gi : it.b->gj].pitch * 8);
        h.v[k].mask = item_mask(it.rq, it.cq, it.b->gi == it.b->gj, shared_is_row);
      }
      return h;
    };
    int n_single = 0;
    for (const DBlock& b : M->blocks)
      if (b.role == r && b.n_items == 1) ++n_single;
    int K = n_single > 40 ? 4 : n_single > 24 ? 3 : n_single > 12 ? 2 : 1;
    std::vector<std::vector<DHalf>> chains;
    {
      std::vector<std::vector<Item>> by_cq(quads);
      for (const DBlock& b : M->blocks)
        if (b.role == r && b.n_items == 1) by_cq[M->groups[b.gj].quad0].push_back(Item{M->groups[b.gi].quad0, M->groups[b.gj].quad0, &b});
      for (const std::vector<Item>& L : by_cq) {
        if (L.empty()) continue;
        const int nch = ((int)L.size() + K - 1) / K;
        for (int c = 0; c < nch; ++c) {
          const size_t lo = L.size() * c / nch, hi = L.size() * (c + 1) / nch;
          chains.push_back(std::vector<DHalf>(1, make_half(std::vector<Item>(L.begin() + lo, L.begin() + hi), false)));
        }
      }
    }
    for (const DBlock& b : M->blocks) {
      if (b.role != r || b.n_items == 1) continue;
      const DGroup &gi = M->groups[b.gi], &gj = M->groups[b.gj];
      const bool shared_is_row = (b.gi == b.gj) || b.tr;  // same group: shared = row quad, element (min, max)
      const DGroup &gs = shared_is_row ? gi : gj, &gv = shared_is_row ? gj : gi;
      std::vector<DHalf> ch;
      for (int qs = 0; qs < gs.n_quads; ++qs) {
        std::vector<Item> its;
        for (int qv = 0; qv < gv.n_quads; ++qv) {
          const int rq = shared_is_row ? gs.quad0 + qs : gv.quad0 + qv;
          const int cq = shared_is_row ? gv.quad0 + qv : gs.quad0 + qs;
          if (b.gi == b.gj && rq > cq) continue;
          its.push_back(Item{rq, cq, &b});
          if ((int)its.size() == PGB_MAX_K) { ch.push_back(make_half(its, shared_is_row)); its.clear(); }
        }
        if (!its.empty()) ch.push_back(make_half(its, shared_is_row));
      }
      chains.push_back(ch);
    }
    // longest chain first onto the least loaded half-warp column
    std::sort(chains.begin(), chains.end(),
              [](const std::vector<DHalf>& x, const std::vector<DHalf>& y) { return x.size() > y.size(); });
    const int nw = std::max(1, std::min<int>(kMaxWarps, ((int)chains.size() + 1) / 2));
    std::vector<std::vector<DHalf>> col(2 * nw);
    for (const std::vector<DHalf>& ch : chains) {
      int best = 0;
      for (int c = 1; c < 2 * nw; ++c)
        if (col[c].size() < col[best].size()) best = c;
      col[best].insert(col[best].end(), ch.begin(), ch.end());
    }
    R.warp0 = (int)M->warps.size();
    R.n_warps = nw;
    R.step0 = (int)M->steps.size();
    for (int w = 0; w < nw; ++w) {
      const size_t ns = std::max(col[2 * w].size(), col[2 * w + 1].size());
      M->warps.push_back(DWarp{(int)M->steps.size(), (int)ns});
      for (size_t q = 0; q < ns; ++q) {
        DStep st{};
        if (q < col[2 * w].size()) st.h[0] = col[2 * w][q];
        if (q < col[2 * w + 1].size()) st.h[1] = col[2 * w + 1][q];
        M->steps.push_back(st);
      }
      R.max_steps = std::max(R.max_steps, (int)ns);
    }
    R.n_steps = (int)M->steps.size() - R.step0;
    max_warps = std::max(max_warps, nw);
    M->roles.push_back(R);
  }

  // ---- every term (and the rhs) with its global record slot: what the records kernel evaluates
  for (int j = 0; j < nt; ++j) M->rterms.push_back(DRoleTerm{j, T[j].gslot0, T[j].idx_base, 0});
  M->rterms.push_back(DRoleTerm{-1, 4 * rhs_quad + rhs_s, rhs_idx, 0});

  // ---- tile rows (shared by both kernels) from the accumulation kernel's shared memory: its
  //      accumulator, kNBuf record buffers of the role's slots, the role's steps
  auto smem_for = [&](int tile) {
    size_t smem = 0;
    for (const DRole& r : M->roles) {
      const size_t need = (size_t)(r.acc_size + 2) * 8 + (size_t)kNBuf * tile_pitch(tile) * r.n_slots * kSlotBytes +
                          (size_t)r.n_steps * sizeof(DStep) + (size_t)r.n_runs * sizeof(DRun) + 8 * kNBuf + 256;
      smem = std::max(smem, need);
    }
    return smem;
  };
  tile = 128;
  while (tile > kMinTile && smem_for(tile) > kSmemMax) tile -= 8;
  for (int t2 = tile; t2 >= 64; t2 -= 8) {
    if (smem_for(t2) + 1024 <= kSmemMax / 2) { tile = t2; break; }
  }
  smem = smem_for(tile);
  const bool fits = smem <= kSmemMax;
  if (final_pass) break;
  if (fits && tile > best_tile) { best_tile = tile; best_budget = budget_d; }
  if (fits && (tile >= 40 || n_roles == 1)) { attempt = 12; continue; }  // good enough: rebuild it as the final plan
  budget_d *= 0.88;
  }  // attempts
  M->tile_rows = tile;
  M->gram_smem = (smem + 127) & ~(size_t)127;
  M->gram_threads = 32 * (max_warps + 1);
  // records kernel: beta, the terms; record slots no term fills (quads of fewer than four slots)
  M->pad_slots.clear();
  {
    std::vector<char> filled(M->n_slots_all, 0);
    for (int j = 0; j < nt; ++j)
      for (int q = 0; q < T[j].nnz; ++q) filled[T[j].gslot0 + q] = 1;
    filled[4 * rhs_quad + rhs_s] = 1;
    for (int q = 0; q < M->n_slots_all; ++q)
      if (!filled[q]) M->pad_slots.push_back(q);
  }
  M->rec_smem = (fixed + (size_t)(nt + 1) * sizeof(DRoleTerm) + M->pad_slots.size() * 4 + 256 + 127) & ~(size_t)127;
  M->rec_ctas = kSMs * 8;
  // ---- CTAs per role proportional to the role's cost per row
  int occ = (int)std::min<size_t>(8, kSmemMax / (M->gram_smem + 1024));
  occ = std::max(1, std::min(occ, 1536 / M->gram_threads));
  const int total = std::max(kSMs * occ, n_roles);
  std::vector<double> cost(n_roles, 0.0);
  double cost_sum = 0.0;
  for (int q = 0; q < n_roles; ++q) {
    const DRole& r = M->roles[q];
    // shared-memory wavefronts per data row vs the longest dependent chain of one warp
    double wf = 0.0, lat = 0.0;
    for (int w = 0; w < r.n_warps; ++w) {
      const DWarp& wp = M->warps[r.warp0 + w];
      double l = 0.0;
      for (int s2 = 0; s2 < wp.n_steps; ++s2) {
        int k2 = 0;
        for (int h = 0; h < 2; ++h)
          for (int k = 0; k < PGB_MAX_K; ++k)
            if (M->steps[wp.step0 + s2].h[h].v[k].mask & 0xffffu) k2 = std::max(k2, k + 1);
        wf += 6.0 * k2 + 3.0;
        l += 40.0 + 10.0 * k2;
      }
      lat = std::max(lat, l);
    }
    cost[q] = std::max(lat, wf * occ) + 0.6 * r.n_slots + 10.0;
    cost_sum += cost[q];
  }
  int assigned = 0;
  for (int q = 0; q < n_roles; ++q) {
    M->roles[q].n_ctas = std::max(1, (int)(total * cost[q] / cost_sum));
    assigned += M->roles[q].n_ctas;
  }
  for (int q = 0; assigned < total; q = (q + 1) % n_roles) { ++M->roles[q].n_ctas; ++assigned; }
  int cta = 0;
  int64_t poff = 0, moff = 0;
  for (DRole& r : M->roles) {
    r.cta0 = cta;
    cta += r.n_ctas;
    r.part_off = poff;
    poff += (int64_t)r.n_ctas * r.acc_size;
    r.map_off = moff;
    moff += r.acc_size;
  }
  M->gram_ctas = cta;
  M->partial_doubles = (poff + 1) & ~(int64_t)1;
  M->map_size = moff;
  M->pass_ctas = kSMs * 8;
  // ---- element map: accumulator element -> row * (m + 1) + col of [G | r], row <= col
  M->elem_map.assign((size_t)moff, -1);
  for (const DBlock& b : M->blocks) {
    const DRole& r = M->roles[b.role];
    const DGroup &gi = M->groups[b.gi], &gj = M->groups[b.gj];
    for (int ri = 0; ri < gi.width; ++ri) {
      const int cr = gcoef[b.gi][ri];
      if (cr == m) continue;  // the rhs is never a row
      for (int ci = (b.gi == b.gj ? ri : 0); ci < gj.width; ++ci) {
        const int cc = gcoef[b.gj][ci];
        const int lo = std::min(cr, cc), hi = std::max(cr, cc);
        const int64_t e = b.tr ? (int64_t)ci * gi.pitch + ri : (int64_t)ri * gj.pitch + ci;
        M->elem_map[(size_t)(r.map_off + b.off + e)] = lo * (m + 1) + hi;
      }
    }
  }
  return PGB_OK;
}

int pgb_plan_upload(pgb_model* M) {
  cudaError_t e = cudaSetDevice(M->device);
  auto up = [&](void** dptr, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dptr, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess && bytes) e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&M->d_steps, M->steps.data(), M->steps.size() * sizeof(DStep));
  up((void**)&M->d_warps, M->warps.data(), M->warps.size() * sizeof(DWarp));
  up((void**)&M->d_rterms, M->rterms.data(), M->rterms.size() * sizeof(DRoleTerm));
  up((void**)&M->d_runs, M->runs.data(), M->runs.size() * sizeof(DRun));
  up((void**)&M->d_pad_slots, M->pad_slots.data(), M->pad_slots.size() * sizeof(int32_t));
  up((void**)&M->d_roles, M->roles.data(), M->roles.size() * sizeof(DRole));
  up((void**)&M->d_elem_map, M->elem_map.data(), M->elem_map.size() * sizeof(int32_t));
  // the device copy of the terms carries the plan fields (group, idx_base, gslot0)
  if (e == cudaSuccess && M->d_terms)
    e = cudaMemcpy(M->d_terms, M->dterms.data(), sizeof(DTerm) * M->dterms.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return pgb_cuda_fail(e, "pgb_plan_upload");
  return PGB_OK;
}

// host-only self check of a plan: every entry of the upper triangle of G and of r must be
// produced by exactly one accumulator element; returns the number of violations
int pgb_plan_coverage_errors(const pgb_model* M) {
  const int m = M->m;
  std::vector<int> hits((size_t)m * (m + 1), 0);
  int errors = 0;
  for (int32_t idx : M->elem_map) {
    if (idx < 0) continue;
    if (idx >= m * (m + 1)) { ++errors; continue; }
    ++hits[idx];
  }
  for (int r = 0; r < m; ++r)
    for (int c = 0; c <= m; ++c) {
      const int want = (c >= r) ? 1 : 0;
      if (hits[(size_t)r * (m + 1) + c] != want) ++errors;
    }
  // every block an item points at must lie inside its role's accumulator, every quad in its record
  for (const DRole& r : M->roles)
    for (int w = 0; w < r.n_warps; ++w) {
      const DWarp& wp = M->warps[r.warp0 + w];
      for (int s = 0; s < wp.n_steps; ++s)
        for (int h = 0; h < 2; ++h) {
          const DHalf& hf = M->steps[wp.step0 + s].h[h];
          if ((int)hf.shared_slot + 4 > r.n_slots) ++errors;
          for (int k = 0; k < PGB_MAX_K; ++k)
            if ((hf.v[k].mask & 0xffffu) && ((int64_t)hf.v[k].off8 / 8 >= r.acc_size || (int)hf.v[k].slot + 4 > r.n_slots))
              ++errors;
        }
    }
  return errors;
}

static void stage_free(pgb_model* M);

void pgb_gram_release(pgb_model* M) {
  stage_free(M);
  for (cudaEvent_t e : M->timing.chunk_ev) cudaEventDestroy(e);
  M->timing.chunk_ev.clear();
  cudaFree(M->d_steps);
  cudaFree(M->d_warps);
  cudaFree(M->d_rterms);
  cudaFree(M->d_runs);
  cudaFree(M->d_pad_slots);
  cudaFree(M->d_roles);
  cudaFree(M->d_elem_map);
  M->d_steps = nullptr; M->d_warps = nullptr; M->d_rterms = nullptr; M->d_runs = nullptr; M->d_pad_slots = nullptr; M->d_roles = nullptr; M->d_elem_map = nullptr;
}

// tiles / record bytes of a chunk of rows; workspace: [partials][scalar partials][records of one chunk]
static int64_t rec_tile_bytes(const pgb_model* M) {
  return (int64_t)M->n_slots_all * tile_pitch(M->tile_rows) * kSlotBytes;
}
// Rows are processed in chunks of whole tiles (bounded record workspace): records kernel, then the
// accumulation kernel continuing from its partials of the previous chunk.  (Running the records
// kernel of the next chunk concurrently on a second stream was measured: it contends for the LSU
// pipe and lengthens the pass by 5 %.)
static int64_t chunk_rows_for(const pgb_model* M, int64_t n) {
  const int64_t T = M->tile_rows;
  if (n <= kChunkRows / 2) return ((n + T - 1) / T) * T;
  const int64_t nch = std::max<int64_t>(2, (n + kChunkRows - 1) / kChunkRows);
  const int64_t rows = (n + nch - 1) / nch;
  return ((rows + T - 1) / T) * T;
}
extern "C" int64_t pgb_cells_ws_bytes(const pgb_model* M, int64_t n, int with_order);  // pgb_cells.cu
int pgb_launch_gram_cells(const pgb_model* M, int32_t mode, const double* X, int64_t n, int64_t ld, const double* y,
                          const float* w, const double* beta, double* out, void* ws, int64_t ws_bytes, int accumulate,
                          int planned, cudaStream_t st);
int pgb_cells_prepare(const pgb_model* M, const double* X, int64_t n, int64_t ld, void* ws, int64_t ws_bytes, cudaStream_t st);
extern "C" int64_t pgb_gram_ws_bytes_streaming(const pgb_model* M, int64_t n);
extern "C" int64_t pgb_gram_ws_bytes(const pgb_model* M, int64_t n) {
  if (!M) return 0;
  if (M->cells.enabled) return pgb_cells_ws_bytes(M, std::max<int64_t>(n, 1), 1);
  const int64_t tiles = chunk_rows_for(M, std::max<int64_t>(n, 1)) / M->tile_rows;
  return (M->partial_doubles + (int64_t)M->rec_ctas * PGB_GS_COUNT) * 8 + tiles * rec_tile_bytes(M) + 1024;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lds_f64(double& v, uint32_t a) {
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
}
__device__ __forceinline__ void lds_u32(uint32_t& v, uint32_t a) {
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
}
// predicated accumulator store: a lane whose predicate is off issues no shared-memory access
__device__ __forceinline__ void sts_f64_if(uint32_t a, double v, uint32_t on) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"(on)
               : "memory");
}

// record loads with a compile-time row offset (records are slot-major: consecutive rows are
// consecutive words, so the offset folds into the instruction)
template <int OFF>
__device__ __forceinline__ void lds_f64_o(double& v, uint32_t a) {
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void lds_u32_o(uint32_t& v, uint32_t a) {
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void lds_v2f64_o(double& x, double& y, uint32_t a) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(a), "n"(OFF));
}
__device__ __forceinline__ void lds_v2u32(uint32_t& x, uint32_t& y, uint32_t a) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
}
__device__ __forceinline__ void st_s16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }

__device__ __forceinline__ double ld_s64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float ld_s32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void st_s64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void st_s32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ld_stterm(uint32_t a, double& lo, double& inv, double& dn, int& kind, int& nint1, int& feature,
                                          int& slot0, int& idx_base, int& coef_offset) {
  int term;
  (void)term;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(lo), "=d"(inv) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(dn) : "r"(a));
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+32];" : "=r"(kind), "=r"(nint1), "=r"(feature), "=r"(slot0) : "r"(a));
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+48];" : "=r"(idx_base), "=r"(term) : "r"(a));
  asm volatile("ld.shared.u32 %0, [%1+56];" : "=r"(coef_offset) : "r"(a));
}

// tile records are slot-major: slot s of tile row r lives at val[s * TP + r] / idx[s * TP + r]
struct EmitRec {
  double* val;     // &val[slot0 * TP + r]
  uint16_t* idx;   // position of the coefficient inside its group
  const double* beta;
  double lp, scale;
  int tp, coef_offset, idx_base;
  __device__ __forceinline__ void operator()(int c, double v) {
    *val = v * scale;
    *idx = (uint16_t)(idx_base + c - coef_offset);
    val += tp;
    idx += tp;
    if (beta) lp = fma(v, beta[c], lp);
  }
};

// One half-step over the rows of a tile, K items in flight, four data rows per trip: per trip two
// 128-bit loads of the shared operand's values, one 64-bit load of its four 16-bit indices, the same
// for each of the K varying operands; then per row K independent read-modify-writes at [varying
// index][shared index] of the items' blocks.  The rows of a trip are updated one after the other
// (they may hit the same element); the operands of the next trip are fetched while the updates of
// the current one are in flight.
template <int K, bool DG>
__device__ __noinline__ void phase_b_half(const DHalf* hp, int tile_n, uint32_t val_u32, uint32_t idx_u32,
                                          uint32_t acc_u32, uint32_t tp, int lt, int lu) {
  const uint32_t ss = (hp->shared_slot + lt) * tp;
  uint32_t a_sv = val_u32 + ss * 8u, a_si = idx_u32 + ss * 2u;
  uint32_t a_vv[K], a_vi[K], base[K], pitch8[K], on[K], dg[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t vs = (hp->v[k].slot + lu) * tp, mk = hp->v[k].mask;
    a_vv[k] = val_u32 + vs * 8u;
    a_vi[k] = idx_u32 + vs * 2u;
    base[k] = acc_u32 + hp->v[k].off8;
    pitch8[k] = hp->v[k].pitch8;
    on[k] = (mk >> (4 * lu + lt)) & 1u;
    dg[k] = (mk >> 16) & 1u;
  }
  double sv[4], vv[K][4];
  uint32_t si[2], vi[K][2];  // four 16-bit indices each
  lds_v2f64_o<0>(sv[0], sv[1], a_sv);
  lds_v2f64_o<16>(sv[2], sv[3], a_sv);
  lds_v2u32(si[0], si[1], a_si);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    lds_v2f64_o<0>(vv[k][0], vv[k][1], a_vv[k]);
    lds_v2f64_o<16>(vv[k][2], vv[k][3], a_vv[k]);
    lds_v2u32(vi[k][0], vi[k][1], a_vi[k]);
  }
  for (int rI = 0; rI < tile_n; rI += 4) {
    uint32_t g[4][K];
    double c_sv[4], c_vv[K][4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      c_sv[h] = sv[h];
      const uint32_t ics = (h & 1) ? (si[h >> 1] >> 16) : (si[h >> 1] & 0xffffu);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        uint32_t ir = (h & 1) ? (vi[k][h >> 1] >> 16) : (vi[k][h >> 1] & 0xffffu), ic = ics;
        if (DG && dg[k] && ir > ic) { const uint32_t t = ir; ir = ic; ic = t; }  // same group: element (min, max)
        g[h][k] = base[k] + ir * pitch8[k] + (ic << 3);
        c_vv[k][h] = vv[k][h];
      }
    }
    a_sv += 32u;
    a_si += 8u;
    lds_v2f64_o<0>(sv[0], sv[1], a_sv);  // the record buffers have four spare rows
    lds_v2f64_o<16>(sv[2], sv[3], a_sv);
    lds_v2u32(si[0], si[1], a_si);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      a_vv[k] += 32u;
      a_vi[k] += 8u;
      lds_v2f64_o<0>(vv[k][0], vv[k][1], a_vv[k]);
      lds_v2f64_o<16>(vv[k][2], vv[k][3], a_vv[k]);
      lds_v2u32(vi[k][0], vi[k][1], a_vi[k]);
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      double x[K];
#pragma unroll
      for (int k = 0; k < K; ++k) lds_f64(x[k], g[h][k]);  // masked lanes read a valid element and do not store
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = fma(c_sv[h], c_vv[k][h], x[k]);
#pragma unroll
      for (int k = 0; k < K; ++k) sts_f64_if(g[h][k], x[k], on[k]);
    }
  }
}

template <int ID>
__device__ __forceinline__ void bar_sync_id(int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive_id(int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
// named barriers 1 .. kNBuf ("record buffer b is free again"); 0 is __syncthreads
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  switch (id) {
    case 1: bar_sync_id<1>(nthreads); break;
    case 2: bar_sync_id<2>(nthreads); break;
    default: bar_sync_id<3>(nthreads); break;
  }
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  switch (id) {
    case 1: bar_arrive_id<1>(nthreads); break;
    case 2: bar_arrive_id<2>(nthreads); break;
    default: bar_arrive_id<3>(nthreads); break;
  }
}

// ----------------------------------------------------------------------------------------
// kernel 1: records.  One thread per data row, no shared-memory staging and no barriers: pass 1
// evaluates the basis of every term for lp; then the IRLS step (mu, W, mask, working response;
// pygam.py:764-777) or the response of _initial_estimate; pass 2 evaluates the basis again and
// writes the row's record = (sqrt(omega) * value, index inside the coefficient group) per slot
// straight to HBM in the slot-major tile layout the accumulation kernel bulk-copies: values
// [tile][slot][pitch], 16-bit indices [tile][slot][pitch].  Consecutive threads own consecutive
// rows, so every feature load and every record store is a contiguous run per warp.  Also the log
// scalars of the iteration.
// ----------------------------------------------------------------------------------------
struct EmitGlobal {
  double* val;    // &val[slot0 * TP + r] of the row's tile
  uint16_t* idx;
  double scale;
  int tp, coef_offset, idx_base;
  __device__ __forceinline__ void operator()(int c, double v) {
    *val = v * scale;
    *idx = (uint16_t)(idx_base + c - coef_offset);
    val += tp;
    idx += tp;
  }
};

__global__ void __launch_bounds__(kRecThreads, 4)
gram_records_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, const DRoleTerm* __restrict__ g_rterms,
                    const int32_t* __restrict__ g_pad_slots, int n_pad, int n_slots, Family fam, int mode,
                    const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                    const float* __restrict__ w, const double* __restrict__ beta, int T, double* __restrict__ rec_val_g,
                    uint16_t* __restrict__ rec_idx_g, double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TP = T + 4, tid = threadIdx.x;
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + PGB_GS_COUNT * 32);
  DRoleTerm* s_rt = reinterpret_cast<DRoleTerm*>(s_terms + n_terms);  // [n_terms + 1], the rhs last
  int32_t* s_pad = reinterpret_cast<int32_t*>(s_rt + n_terms + 1);
  for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  for (int i = tid; i <= n_terms; i += blockDim.x) s_rt[i] = g_rterms[i];
  for (int i = tid; i < n_pad; i += blockDim.x) s_pad[i] = g_pad_slots[i];
  __syncthreads();

  double sc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) sc[s] = 0.0;
  const int64_t n_tiles = (n + T - 1) / T;
  const int64_t n_pad_rows = n_tiles * T;  // the rows that fill up the last tile get all-zero records
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n_pad_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = i / T;
    const int r = (int)(i - tile * T);
    double* rv = rec_val_g + (size_t)tile * n_slots * TP + r;
    uint16_t* ri = rec_idx_g + (size_t)tile * n_slots * TP + r;
    if (i >= n) {
      for (int s = 0; s < n_slots; ++s) { rv[(size_t)s * TP] = 0.0; ri[(size_t)s * TP] = 0; }
      continue;
    }
    const double yi = __ldg(y + i);
    double sw, z;
    if (mode == PGB_MODE_IRLS) {
      EmitLp e{s_beta, 0.0};
      for (int q = 0; q < n_terms; ++q) eval_term(s_terms[q], X, ld, i, e);
      const IrlsRow ir = irls_row(fam, e.lp, yi, weight_inv(w, i));
      sw = ir.sw;
      z = ir.z;
      if (ir.keep) {
        sc[PGB_GS_NUSED] += 1.0;
        sc[PGB_GS_DEV] += dist_dev(fam, yi, ir.mu);
        sc[PGB_GS_NCORRECT] += (yi == ((ir.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      sw = 1.0;
      z = init_response(fam, yi);
      sc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(z)) sc[PGB_GS_NONFINITE] += 1.0;
    }
    sc[PGB_GS_NROWS] += 1.0;
    for (int q = 0; q < n_terms; ++q) {
      const DRoleTerm rt = s_rt[q];
      const DTerm& d = s_terms[q];
      EmitGlobal e{rv + (size_t)rt.slot0 * TP, ri + (size_t)rt.slot0 * TP, sw, TP, d.coef_offset, rt.idx_base};
      eval_term(d, X, ld, i, e);
    }
    {
      const DRoleTerm rt = s_rt[n_terms];  // the rhs column: sqrt(omega) * z
      rv[(size_t)rt.slot0 * TP] = sw * z;
      ri[(size_t)rt.slot0 * TP] = (uint16_t)rt.idx_base;
    }
    for (int q = 0; q < n_pad; ++q) {  // slots no term fills: masked in the accumulation, but read
      rv[(size_t)s_pad[q] * TP] = 0.0;
      ri[(size_t)s_pad[q] * TP] = 0;
    }
  }
  block_reduce_store<PGB_GS_COUNT>(sc, s_red, scal_part + (int64_t)blockIdx.x * PGB_GS_COUNT);
}

// ----------------------------------------------------------------------------------------
// kernel 2: accumulation.  G = sum over rows of record (x) record: the same arithmetic as the
// reference's W.B (pygam.py:782) before its QR.  A CTA works for one role: it keeps the role's
// blocks of G in shared memory and walks its slice of the tiles.  Warps 0 .. n_warps-1 are
// consumers (phase_b_half); lane 0 of the next warp is the producer: TMA bulk copies of the role's
// record slots of a tile (one per run of slots, values and indices) into a ring of kNBuf buffers,
// completion on mbarriers; the consumers hand a buffer back through a named barrier.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * (kMaxWarps + 1))
gram_accum_kernel(const DStep* __restrict__ g_steps, const DWarp* __restrict__ g_warps, const DRun* __restrict__ g_runs,
                  const DRole* __restrict__ g_roles, int n_roles, const double* __restrict__ rec_val_g,
                  const uint16_t* __restrict__ rec_idx_g, int64_t n, int T, int n_slots_all, int dbg, int resume,
                  double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int ridx = 0;
  while (ridx + 1 < n_roles && (int)blockIdx.x >= g_roles[ridx + 1].cta0) ++ridx;
  const DRole role = g_roles[ridx];
  const int cta_in_role = blockIdx.x - role.cta0;
  const int S = role.n_slots, TP = T + 4;

  double* acc = reinterpret_cast<double*>(smem_raw);
  double* rec_val = acc + role.acc_size;                                          // [kNBuf][S][TP]
  uint16_t* rec_idx = reinterpret_cast<uint16_t*>(rec_val + (size_t)kNBuf * S * TP);  // [kNBuf][S][TP]
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(rec_idx + (size_t)kNBuf * S * TP);
  DStep* steps = reinterpret_cast<DStep*>(mbar + kNBuf);
  DRun* runs = reinterpret_cast<DRun*>(steps + role.n_steps);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    // later row chunks of one pass continue from this CTA's partial of the previous chunk
    const double* prev = partials + role.part_off + (int64_t)cta_in_role * role.acc_size;
    for (int64_t i = tid; i < role.acc_size; i += blockDim.x) acc[i] = resume ? prev[i] : 0.0;
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_steps + role.step0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(steps);
    for (int i = tid; i < role.n_steps * (int)(sizeof(DStep) / 4); i += blockDim.x) dst[i] = src[i];
    for (int i = tid; i < role.n_runs; i += blockDim.x) runs[i] = g_runs[role.run0 + i];
  }
  if (tid == 0) {
    for (int q = 0; q < kNBuf; ++q) mbar_init(smem_u32(mbar + q), 1);
    fence_mbar_init();
  }
  __syncthreads();

  // tiles of this CTA: a contiguous slice
  const int64_t n_tiles = (n + T - 1) / T;
  const int64_t tiles_lo = n_tiles * cta_in_role / role.n_ctas;
  const int64_t tiles_hi = n_tiles * (cta_in_role + 1) / role.n_ctas;
  const int nbar = 32 * (role.n_warps + 1);  // consumers + the producer warp

  if (warp == role.n_warps) {
    // ---------------- producer ----------------
    const uint32_t bytes = (uint32_t)((size_t)S * TP * kSlotBytes);
    int b = 0;
    for (int64_t tile = tiles_lo; tile < tiles_hi; ++tile, b = (b + 1 == kNBuf) ? 0 : b + 1) {
      if (tile >= tiles_lo + kNBuf) bar_sync(1 + b, nbar);  // the consumers are done with this buffer
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the consumers' generic reads vs the async writes
        const uint32_t bar = smem_u32(mbar + b);
        mbar_expect_tx(bar, bytes);
        const double* gv = rec_val_g + (size_t)tile * n_slots_all * TP;
        const uint16_t* gi = rec_idx_g + (size_t)tile * n_slots_all * TP;
        double* sv = rec_val + (size_t)b * S * TP;
        uint16_t* si = rec_idx + (size_t)b * S * TP;
        for (int q = 0; q < role.n_runs; ++q) {
          const DRun r = runs[q];
          bulk_g2s(smem_u32(sv + (size_t)r.lslot0 * TP), gv + (size_t)r.gslot0 * TP, (uint32_t)(r.n_slots * TP * 8), bar);
          bulk_g2s(smem_u32(si + (size_t)r.lslot0 * TP), gi + (size_t)r.gslot0 * TP, (uint32_t)(r.n_slots * TP * 2), bar);
        }
      }
      __syncwarp();
    }
  } else if (warp < role.n_warps) {
    // ---------------- consumers: every warp walks the tile once per step ----------------
    const DWarp wp = g_warps[role.warp0 + warp];
    const int half = lane >> 4, lu = (lane >> 2) & 3, lt = lane & 3;
    const uint32_t acc_u32 = smem_u32(acc);
    uint32_t phases = 0u;  // bit b: parity the next wait on buffer b expects
    int b = 0;
    for (int64_t tile = tiles_lo; tile < tiles_hi; ++tile, b = (b + 1 == kNBuf) ? 0 : b + 1) {
      mbar_wait(smem_u32(mbar + b), (phases >> b) & 1u);
      phases ^= 1u << b;
      const int tile_n = (int)min((int64_t)T, n - tile * T);
      const uint32_t val_u32 = smem_u32(rec_val + (size_t)b * S * TP), idx_u32 = smem_u32(rec_idx + (size_t)b * S * TP);
      for (int s = 0; s < ((dbg & 2) ? 0 : wp.n_steps); ++s) {  // dbg 2: no accumulation (timing experiments)
        const DHalf* hp = &steps[wp.step0 - role.step0 + s].h[half];
        // items in flight: the larger of the two half-steps (warp-uniform trip count)
        int kk = 0, dg = 0;
#pragma unroll
        for (int k = 0; k < PGB_MAX_K; ++k) {
          const uint32_t mk = hp->v[k].mask;
          if (mk & 0xffffu) kk = k + 1;
          dg |= (int)(mk >> 16);
        }
        kk = max(kk, __shfl_xor_sync(0xffffffffu, kk, 16));
        dg |= __shfl_xor_sync(0xffffffffu, dg, 16);
        if (dg) {
          switch (kk) {
            case 1: phase_b_half<1, true>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 2: phase_b_half<2, true>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 3: phase_b_half<3, true>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 4: phase_b_half<4, true>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            default: break;
          }
        } else {
          switch (kk) {
            case 1: phase_b_half<1, false>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 2: phase_b_half<2, false>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 3: phase_b_half<3, false>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            case 4: phase_b_half<4, false>(hp, tile_n, val_u32, idx_u32, acc_u32, TP, lt, lu); break;
            default: break;
          }
        }
      }
      if (tile + kNBuf < tiles_hi) bar_arrive(1 + b, nbar);
    }
  }
  __syncthreads();
  // ---------------- write the partial ----------------
  double* dst = partials + role.part_off + (int64_t)cta_in_role * role.acc_size;
  for (int64_t i = tid; i < role.acc_size; i += blockDim.x) dst[i] = acc[i];
}

// out = [G (m*m, symmetric) | r (m)]: sum the CTA partials of each role in order, mirror
__global__ void gram_reduce_kernel(const double* __restrict__ partials, const DRole* __restrict__ roles,
                                   int n_roles, const int32_t* __restrict__ elem_map, int64_t map_size, int m,
                                   double* __restrict__ out, int accumulate) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= map_size) return;
  const int idx = elem_map[e];
  if (idx < 0) return;
  int r = 0;
  while (r + 1 < n_roles && e >= roles[r + 1].map_off) ++r;
  const DRole role = roles[r];
  const double* p = partials + role.part_off + (e - role.map_off);
  double v = 0.0;
  for (int c = 0; c < role.n_ctas; ++c) v += p[(int64_t)c * role.acc_size];
  const int row = idx / (m + 1), col = idx - row * (m + 1);
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
  const int planned = (mode & PGB_MODE_PLANNED) ? 1 : 0;
  mode &= ~PGB_MODE_PLANNED;
  if (!M || !X || !y || !out || !ws || n < 1 || ld < n || (mode == PGB_MODE_IRLS && !beta) ||
      (mode != PGB_MODE_IRLS && mode != PGB_MODE_INIT)) {
    pgb_set_error("pgb_gram_pass: bad arguments");
    return PGB_EINVAL;
  }
  const int64_t ws_need = planned ? pgb_gram_ws_bytes(M, n) : pgb_gram_ws_bytes_streaming(M, n);
  if (ws_bytes < ws_need) {
    pgb_set_error("pgb_gram_pass: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)ws_need);
    return PGB_EINVAL;
  }
  if (M->cells.enabled) return pgb_launch_gram_cells(M, mode, X, n, ld, y, w, beta, out, ws, ws_bytes, accumulate, planned, st);
  const int n_roles = (int)M->roles.size();
  const int T = M->tile_rows, TP = tile_pitch(T);
  pgb_model* Mm = const_cast<pgb_model*>(M);
  double* partials = (double*)ws;
  double* scal_part = partials + M->partial_doubles;
  // records of one chunk: values then indices, 256-byte aligned
  const int64_t chunk_rows = chunk_rows_for(M, n);
  const int64_t chunk_tiles = chunk_rows / T;
  uintptr_t rp = (uintptr_t)(scal_part + (int64_t)M->rec_ctas * PGB_GS_COUNT);
  rp = (rp + 255) & ~(uintptr_t)255;
  double* rec_val = (double*)rp;
  uint16_t* rec_idx = (uint16_t*)(rec_val + chunk_tiles * M->n_slots_all * TP);
  // the attribute is per device and cheap to set: no caching (a thread may drive engines on several GPUs)
  PGB_CUDA(cudaFuncSetAttribute(gram_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
  PGB_CUDA(cudaFuncSetAttribute(gram_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
  const int dbg = 0;
  pgb_model::Timing& tm = Mm->timing;
  const int n_chunks = (int)((n + chunk_rows - 1) / chunk_rows);
  if (tm.enabled) {
    while ((int)tm.chunk_ev.size() < 3 * n_chunks + 2) {
      cudaEvent_t e;
      PGB_CUDA(cudaEventCreate(&e));
      tm.chunk_ev.push_back(e);
    }
    PGB_CUDA(cudaEventRecord(tm.chunk_ev[0], st));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t r0 = (int64_t)c * chunk_rows, nr = std::min(chunk_rows, n - r0);
    const int64_t tiles = (nr + T - 1) / T;
    const int grid1 = (int)std::min<int64_t>((tiles * T + kRecThreads - 1) / kRecThreads, M->rec_ctas);
    if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.chunk_ev[2 + 3 * c], st));
    gram_records_kernel<<<grid1, kRecThreads, M->rec_smem, st>>>(M->d_terms, M->n_terms, M->m, M->d_rterms, M->d_pad_slots,
                                                                   (int)M->pad_slots.size(), M->n_slots_all, M->fam, mode, X + r0,
                                                                   nr, ld, y + r0, w ? w + r0 : nullptr, beta, T, rec_val, rec_idx,
                                                                   scal_part);
    PGB_CUDA(cudaGetLastError());
    if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.chunk_ev[3 + 3 * c], st));
    gram_accum_kernel<<<M->gram_ctas, M->gram_threads, M->gram_smem, st>>>(M->d_steps, M->d_warps, M->d_runs, M->d_roles, n_roles,
                                                                              rec_val, rec_idx, nr, T, M->n_slots_all, dbg,
                                                                              c > 0 ? 1 : 0, partials);
    PGB_CUDA(cudaGetLastError());
    if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.chunk_ev[4 + 3 * c], st));
    scalar_reduce_kernel<<<1, 32 * PGB_GS_COUNT, 0, st>>>(scal_part, grid1, PGB_GS_COUNT, out + (int64_t)M->m * M->m + M->m,
                                           (accumulate || c > 0) ? 1 : 0);
    pgb_count_launch(3);
  }
  const int rb = 256;
  gram_reduce_kernel<<<(unsigned)((M->map_size + rb - 1) / rb), rb, 0, st>>>(partials, M->d_roles, n_roles, M->d_elem_map,
                                                                              M->map_size, M->m, out, accumulate);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  if (tm.enabled) {
    PGB_CUDA(cudaEventRecord(tm.chunk_ev[1], st));
    tm.pending = true;
    tm.n_chunks = n_chunks;
  }
  return PGB_OK;
}

// optional timing of the Gram pass with CUDA events on the launching stream: the two kernels (summed
// over the row chunks) and the whole pass; used by bench.py for the roofline figure
extern "C" int pgb_gram_timing(pgb_model* M, int32_t enable) {
  if (!M) { pgb_set_error("pgb_gram_timing: null model"); return PGB_EINVAL; }
  M->timing.enabled = enable != 0;
  M->timing.pending = false;
  return PGB_OK;
}
extern "C" int pgb_gram_last_ms(pgb_model* M, float* records_ms, float* accum_ms, float* total_ms) {
  if (!M || !records_ms || !accum_ms || !total_ms || !M->timing.pending) { pgb_set_error("pgb_gram_last_ms: no timed pass"); return PGB_EINVAL; }
  pgb_model::Timing& tm = M->timing;
  *records_ms = 0.f;
  *accum_ms = 0.f;
  PGB_CUDA(cudaEventSynchronize(tm.chunk_ev[1]));
  PGB_CUDA(cudaEventElapsedTime(total_ms, tm.chunk_ev[0], tm.chunk_ev[1]));
  for (int c = 0; c < tm.n_chunks; ++c) {
    float a = 0.f, b = 0.f;
    PGB_CUDA(cudaEventElapsedTime(&a, tm.chunk_ev[2 + 3 * c], tm.chunk_ev[3 + 3 * c]));
    PGB_CUDA(cudaEventElapsedTime(&b, tm.chunk_ev[3 + 3 * c], tm.chunk_ev[4 + 3 * c]));
    *records_ms += a;
    *accum_ms += b;
  }
  return PGB_OK;
}

// workspace of passes that never use pgb_gram_prepare (rows streamed through once, or not enough memory for
// the stored order: 2 bytes per row and term pair)
extern "C" int64_t pgb_gram_ws_bytes_streaming(const pgb_model* M, int64_t n) {
  if (!M) return 0;
  if (M->cells.enabled) return pgb_cells_ws_bytes(M, std::max<int64_t>(n, 1), 0);
  return pgb_gram_ws_bytes(M, n);
}

// The knot cells of a row, and therefore the cell order of the rows, do not depend on beta: a fit prepares
// its resident shard once and passes PGB_MODE_PLANNED afterwards.  No-op for models on the general path.
extern "C" int pgb_gram_prepare(const pgb_model* M, const double* X, int64_t n, int64_t ld, void* ws, int64_t ws_bytes,
                                void* stream) {
  if (!M || !X || !ws || n < 1 || ld < n) { pgb_set_error("pgb_gram_prepare: bad arguments"); return PGB_EINVAL; }
  if (!M->cells.enabled) return PGB_OK;
  return pgb_cells_prepare(M, X, n, ld, ws, ws_bytes, (cudaStream_t)stream);
}

int pgb_cells_refine(const pgb_model* M, int64_t n, void* ws, int64_t ws_bytes, cudaStream_t st);  // pgb_cells.cu
extern "C" int pgb_gram_refine(const pgb_model* M, int64_t n, void* ws, int64_t ws_bytes, void* stream) {
  if (!M || !ws || n < 1) { pgb_set_error("pgb_gram_refine: bad arguments"); return PGB_EINVAL; }
  if (!M->cells.enabled) return PGB_OK;
  return pgb_cells_refine(M, n, ws, ws_bytes, (cudaStream_t)stream);
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
  S.ws_bytes = pgb_gram_ws_bytes_streaming(M, chunk_rows);
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
