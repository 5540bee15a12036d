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

static const int kMaxWarps = 12;      // consumer (accumulating) warps per CTA of the Gram pass
#ifndef PGB_STAGE_U
#define PGB_STAGE_U 2
#endif
static const int kMaxProducers = 4;
static const int kStages = 4;        // raw-column stage buffers of the producers (fetch distance kStages - 1 tiles)  // producer (record staging) warps per CTA
static const int kSlotBytes = 12;     // record slot: double value + uint32 index (two arrays)

// ----------------------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------------------
static const int kMinTile = 16;
// row pitch of the slot-major tile records: odd (the 4 slots of a quad start in different banks) and
// two spare rows (phase B prefetches one row pair ahead)
static int tile_pitch(int tile) { return (tile + 2) | 1; }

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

  // ---- blocks, packed into roles in (gi, gj) order
  const size_t fixed = role_fixed_smem(M);
  const size_t rec_row_all = (size_t)(quads * 4) * kSlotBytes;  // bytes per tile row when a role needs every quad
  if (fixed + 2 * kMinTile * rec_row_all + 8192 > kSmemMax) {
    pgb_set_error("model too wide for the Gram pass: k=%d non-zeros per row", M->k);
    return PGB_EUNSUPPORTED;
  }
  const size_t stage_min = (size_t)kStages * (M->n_features + 2) * kMinTile * 8 + (size_t)4 * kStages * kMinTile;  // raw-column stage buffers
  const int64_t acc_budget = (int64_t)((kSmemMax - fixed - 2 * (kMinTile + 3) * rec_row_all - stage_min - 6144) / 8);  // doubles
  M->blocks.clear();
  int64_t total_acc = 0;
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
      if (sz > acc_budget) {
        pgb_set_error("a Gram block of %d x %d coefficients does not fit in shared memory", M->groups[gi].width,
                      M->groups[gj].width);
        return PGB_EUNSUPPORTED;
      }
      total_acc += sz;
      M->blocks.push_back(b);
    }
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
  const int n_roles = (int)role_acc.size();

  // ---- per role: record layout, terms to evaluate, items -> steps -> warps
  M->roles.clear();
  M->rterms.clear();
  M->steps.clear();
  M->warps.clear();
  int max_warps = 1;
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
    R.term0 = (int)M->rterms.size();
    for (int j = 0; j < nt; ++j) {
      if (!gneed[T[j].group]) continue;
      const int gq = T[j].gslot0 >> 2;
      M->rterms.push_back(DRoleTerm{j, 4 * lquad[gq] + (T[j].gslot0 & 3), T[j].idx_base, 0});
    }
    R.rhs_slot = -1;
    if (gneed[rhs_group]) {
      R.rhs_slot = 4 * lquad[rhs_quad] + rhs_s;
      M->rterms.push_back(DRoleTerm{-1, R.rhs_slot, rhs_idx, 0});
    }
    R.n_rterms = (int)M->rterms.size() - R.term0;
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
    int K = n_single > 60 ? 4 : n_single > 40 ? 3 : n_single > 12 ? 2 : 1;
    if (const char* ek = getenv("PGB_GRAM_K")) K = std::max(1, std::min(PGB_MAX_K, atoi(ek)));
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

  // ---- tile rows, shared memory, threads: the largest tile that fits; if a tile of >= 64 rows
  //      also fits twice per SM, take that one (two resident CTAs hide each other's phase A)
  // records are slot-major ([slot][tile_pitch]); split roles double-buffer them so that staging the
  // next tile overlaps the accumulation of the current one
  auto smem_for = [&](int tile) {
    size_t smem = 0;
    const int nbuf = 2;
    for (const DRole& r : M->roles) {
      size_t need = (size_t)(r.acc_size + 2) * 8 + (size_t)nbuf * tile_pitch(tile) * r.n_slots * kSlotBytes + (size_t)tile * 8 + fixed +
                    (size_t)r.n_steps * sizeof(DStep) + (size_t)r.n_rterms * 64 +
                    (size_t)kStages * (M->n_features + 2) * tile * 8 + (size_t)kStages * tile * 4 + 8 * kStages + 64;
      if (n_roles == 1) need += (size_t)tile * r.n_rterms * 8;  // lp_part of the inline IRLS step
      smem = std::max(smem, need);
    }
    return smem;
  };
  int tile = 128;
  if (const char* et = getenv("PGB_GRAM_TILE")) tile = std::max(kMinTile, atoi(et) & ~7);
  while (tile > kMinTile && smem_for(tile) > kSmemMax) tile -= 8;
  for (int t2 = tile; t2 >= 64; t2 -= 8) {
    if (smem_for(t2) + 1024 <= kSmemMax / 2) { tile = t2; break; }
  }
  size_t smem = smem_for(tile);
  if (smem > kSmemMax) {
    pgb_set_error("Gram pass does not fit in shared memory (%zu bytes)", smem);
    return PGB_EUNSUPPORTED;
  }
  M->tile_rows = tile;
  M->gram_smem = (smem + 127) & ~(size_t)127;
  int prod = kMaxProducers;
  if (const char* ep = getenv("PGB_GRAM_PW")) prod = std::max(1, std::min(kMaxProducers, atoi(ep)));
  M->gram_threads = 32 * (max_warps + prod);
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
  M->partial_doubles = (poff + 1) & ~(int64_t)1;  // keeps the (sqrt(omega), z) pairs after it 16-byte aligned
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
  if (M->timing.ev[0]) { cudaEventDestroy(M->timing.ev[0]); cudaEventDestroy(M->timing.ev[1]); }
  cudaFree(M->d_steps);
  cudaFree(M->d_warps);
  cudaFree(M->d_rterms);
  cudaFree(M->d_roles);
  cudaFree(M->d_elem_map);
  M->d_steps = nullptr; M->d_warps = nullptr; M->d_rterms = nullptr; M->d_roles = nullptr; M->d_elem_map = nullptr;
}

static bool gram_inline(const pgb_model* M) { return M->roles.size() == 1 && M->roles[0].needs_all_terms; }

// workspace: [partials][scalar partials][(omega, z) per row when split]
static int64_t gram_fixed_ws_doubles(const pgb_model* M) {
  return M->partial_doubles + (int64_t)(M->gram_ctas + M->pass_ctas) * PGB_GS_COUNT;
}
extern "C" int64_t pgb_gram_ws_bytes(const pgb_model* M, int64_t n) {
  if (!M) return 0;
  int64_t d = gram_fixed_ws_doubles(M);
  if (!gram_inline(M)) d += 2 * n;
  return d * 8 + 256;
}

// ----------------------------------------------------------------------------------------
// per-row IRLS quantities when the Gram matrix is split over several roles: every role's
// CTAs read (omega, z) instead of re-evaluating the full linear predictor
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPassThreads)
irls_rows_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, Family fam, int mode,
                 const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                 const float* __restrict__ w, const double* __restrict__ beta,
                 double* __restrict__ wz, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_beta = reinterpret_cast<double*>(smem_raw);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_beta + ((m + 1) & ~1));
  double* s_red = reinterpret_cast<double*>(s_terms + n_terms);
  for (int i = threadIdx.x; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();
  double acc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) acc[s] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yi = __ldg(y + i);
    double om, z;
    if (mode == PGB_MODE_IRLS) {
      EmitLp e{s_beta, 0.0};
      for (int t = 0; t < n_terms; ++t) eval_term(s_terms[t], X, ld, i, e);
      const IrlsRow r = irls_row(fam, e.lp, yi, weight_inv(w, i));
      om = r.sw;
      z = r.z;
      if (r.keep) {
        acc[PGB_GS_NUSED] += 1.0;
        acc[PGB_GS_DEV] += dist_dev(fam, yi, r.mu);
        acc[PGB_GS_NCORRECT] += (yi == ((r.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
      }
    } else {
      om = 1.0;
      z = init_response(fam, yi);
      acc[PGB_GS_NUSED] += 1.0;
      if (!isfinite(z)) acc[PGB_GS_NONFINITE] += 1.0;
    }
    acc[PGB_GS_NROWS] += 1.0;
    reinterpret_cast<double2*>(wz)[i] = make_double2(om, om * z);  // (sqrt(omega), sqrt(omega) z)
  }
  block_reduce_store<PGB_GS_COUNT>(acc, s_red, part + (int64_t)blockIdx.x * PGB_GS_COUNT);
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
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(lo), "=d"(inv) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(dn) : "r"(a));
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+32];" : "=r"(kind), "=r"(nint1), "=r"(feature), "=r"(slot0) : "r"(a));
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+48];" : "=r"(idx_base), "=r"(term) : "r"(a));
  asm volatile("ld.shared.u32 %0, [%1+56];" : "=r"(coef_offset) : "r"(a));
}

// tile records are slot-major: slot s of tile row r lives at val[s * TP + r] / idx[s * TP + r]
struct EmitRec {
  double* val;    // &val[slot0 * TP + r]
  uint32_t* idx;  // position of the coefficient inside its group
  const double* beta;
  double lp, scale;
  int tp, coef_offset, idx_base;
  __device__ __forceinline__ void operator()(int c, double v) {
    *val = v * scale;
    *idx = (uint32_t)(idx_base + c - coef_offset);
    val += tp;
    idx += tp;
    if (beta) lp = fma(v, beta[c], lp);
  }
};

// One half-step of phase B over the rows of a tile, K items in flight, two data rows per trip:
// per data row one load of the shared operand (value, index) and K loads of the varying operands,
// then K independent read-modify-writes at [varying index][shared index] of the items' blocks.
// The two rows of a trip are updated one after the other (they may hit the same element); the
// operands of the next trip are fetched while the updates of the current one are in flight.
template <int K, bool DG>
__device__ __noinline__ void phase_b_half(const DHalf* hp, int tile_n, uint32_t val_u32,
                                             uint32_t idx_u32, uint32_t acc_u32, uint32_t tp, int lt, int lu) {
  const uint32_t ss = (hp->shared_slot + lt) * tp;
  uint32_t a_sv = val_u32 + ss * 8u, a_si = idx_u32 + ss * 4u;
  uint32_t a_vv[K], a_vi[K], base[K], pitch8[K], on[K], dg[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t vs = (hp->v[k].slot + lu) * tp, mk = hp->v[k].mask;
    a_vv[k] = val_u32 + vs * 8u;
    a_vi[k] = idx_u32 + vs * 4u;
    base[k] = acc_u32 + hp->v[k].off8;
    pitch8[k] = hp->v[k].pitch8;
    on[k] = (mk >> (4 * lu + lt)) & 1u;
    dg[k] = (mk >> 16) & 1u;
  }
  double sv[2], vv[2][K];
  uint32_t si[2], vi[2][K];
  lds_f64_o<0>(sv[0], a_sv);
  lds_f64_o<8>(sv[1], a_sv);
  lds_u32_o<0>(si[0], a_si);
  lds_u32_o<4>(si[1], a_si);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    lds_f64_o<0>(vv[0][k], a_vv[k]);
    lds_f64_o<8>(vv[1][k], a_vv[k]);
    lds_u32_o<0>(vi[0][k], a_vi[k]);
    lds_u32_o<4>(vi[1][k], a_vi[k]);
  }
  for (int rI = 0; rI < tile_n; rI += 2) {
    uint32_t g[2][K];
    double c_sv[2], c_vv[2][K];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      c_sv[h] = sv[h];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        uint32_t ir = vi[h][k], ic = si[h];
        if (DG && dg[k] && ir > ic) { const uint32_t t = ir; ir = ic; ic = t; }  // same group: element (min, max)
        g[h][k] = base[k] + ir * pitch8[k] + (ic << 3);
        c_vv[h][k] = vv[h][k];
      }
    }
    a_sv += 16u;
    a_si += 8u;
    lds_f64_o<0>(sv[0], a_sv);  // the record buffers have two spare rows
    lds_f64_o<8>(sv[1], a_sv);
    lds_u32_o<0>(si[0], a_si);
    lds_u32_o<4>(si[1], a_si);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      a_vv[k] += 16u;
      a_vi[k] += 8u;
      lds_f64_o<0>(vv[0][k], a_vv[k]);
      lds_f64_o<8>(vv[1][k], a_vv[k]);
      lds_u32_o<0>(vi[0][k], a_vi[k]);
      lds_u32_o<4>(vi[1][k], a_vi[k]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double x[K];
#pragma unroll
      for (int k = 0; k < K; ++k) lds_f64(x[k], g[h][k]);  // masked lanes read a valid element and do not store
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = fma(c_sv[h], c_vv[h][k], x[k]);
#pragma unroll
      for (int k = 0; k < K; ++k) sts_f64_if(g[h][k], x[k], on[k]);
    }
  }
}

struct StTerm;
struct GramSmem {
  DStep* steps;    // the role's steps
  StTerm* rt;      // the role's terms, compact
  double* xs[kStages];   // raw columns of the tile being staged and of the next ones
  float* ws[kStages];
  unsigned long long* mbar;
  double* acc;
  double* rec_val[2];
  uint32_t* rec_idx[2];
  double* rec_w;
  double* lp_part;
  double* s_beta;
  double* s_red;
  DTerm* s_terms;
};

// named barriers (0 is __syncthreads): the producer and the consumer warps of a CTA hand the two
// record buffers back and forth; arrive does not wait
template <int ID>
__device__ __forceinline__ void bar_sync_id(int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive_id(int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
enum { kBarFull = 1, kBarEmpty = 3, kBarProd = 5 };
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  switch (id) {
    case 1: bar_sync_id<1>(nthreads); break;
    case 2: bar_sync_id<2>(nthreads); break;
    case 3: bar_sync_id<3>(nthreads); break;
    case 4: bar_sync_id<4>(nthreads); break;
    default: bar_sync_id<5>(nthreads); break;
  }
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  switch (id) {
    case 1: bar_arrive_id<1>(nthreads); break;
    case 2: bar_arrive_id<2>(nthreads); break;
    case 3: bar_arrive_id<3>(nthreads); break;
    default: bar_arrive_id<4>(nthreads); break;
  }
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

// compact per-role-term descriptor the producers work from (built in shared memory at kernel
// start): everything the common record kinds need without touching the full DTerm
enum { kStCubic = 0, kStIntercept = 1, kStLinear = 2, kStRhs = 3, kStGeneric = 4 };
struct __align__(16) StTerm {
  double lo, inv, dn, pad0;  // cubic spline: edge knot, 1 / span, number of intervals
  int32_t kind, nint1, feature, slot0;
  int32_t idx_base, term, coef_offset, pad1;
};

struct StageArgs {
  const StTerm* rt;
  const DRoleTerm* g_rterms;
  const DTerm* s_terms;
  const double* X;
  const double* y;
  const float* w;
  const double* wz;
  const double* s_beta;
  double* lp_part;
  double* rec_w;
  double* xs[kStages];   // raw columns of a tile: [F][T] features, then [T] y (inline) or [2T] (sqrt(omega), sqrt(omega) z) (split)
  float* ws[kStages];    // [T] prior weights
  unsigned long long* mbar;  // [kStages] completion barriers of the stage buffers
  int64_t ld, n;
  Family fam;
  int mode, inline_irls, T, TP, F, bulk_ok;
};

// Fetch one tile's raw columns into stage buffer b: TMA bulk copies (one per column, issued by
// producer thread 0, completing on the buffer's mbarrier) when everything is 16-byte aligned,
// plain loads by all producer threads for unaligned inputs and the ragged last tile.
__device__ __forceinline__ void stage_fetch(const StageArgs& A, int64_t row0, int rows, int b, int ptid, int pn) {
  double* xs = A.xs[b];
  const int T = A.T;
  double* tail = xs + (size_t)A.F * T;
  const uint32_t bar = smem_u32(A.mbar + b);
  if (A.bulk_ok && (rows & 3) == 0) {
    if (ptid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this buffer vs the async writes
      const uint32_t bytes = (uint32_t)(A.F * rows * 8 + (A.inline_irls ? rows * 8 + (A.w ? rows * 4 : 0) : rows * 16));
      mbar_expect_tx(bar, bytes);
      for (int f = 0; f < A.F; ++f) bulk_g2s(smem_u32(xs + (size_t)f * T), A.X + (int64_t)f * A.ld + row0, rows * 8, bar);
      if (A.inline_irls) {
        bulk_g2s(smem_u32(tail), A.y + row0, rows * 8, bar);
        if (A.w) bulk_g2s(smem_u32(A.ws[b]), A.w + row0, rows * 4, bar);
      } else {
        bulk_g2s(smem_u32(tail), A.wz + 2 * row0, rows * 16, bar);
      }
    }
    return;
  }
  for (int e = ptid; e < A.F * rows; e += pn) {
    const int f = e / rows, r = e - f * rows;
    xs[(size_t)f * T + r] = __ldg(A.X + (int64_t)f * A.ld + row0 + r);
  }
  if (A.inline_irls) {
    for (int r = ptid; r < rows; r += pn) tail[r] = __ldg(A.y + row0 + r);
    if (A.w)
      for (int r = ptid; r < rows; r += pn) A.ws[b][r] = __ldg(A.w + row0 + r);
  } else {
    for (int r = ptid; r < 2 * rows; r += pn) tail[r] = __ldg(A.wz + 2 * row0 + r);
  }
  if (ptid == 0) mbar_arrive(bar);  // visibility of the plain stores: the producers' barrier of the consuming trip
}

// Phase A of one tile, run by the producer warps (ptid = thread index among them, pn = their
// count) on the raw columns staged in shared memory: one thread per (row, term), four items per
// trip -- all loads and the arithmetic of the four items first (independent chains the scheduler can
// interleave), then their record stores.  Split roles read (sqrt(omega), sqrt(omega) z) written by
// irls_rows_kernel; a single role evaluates the IRLS step itself: term contributions to lp, then
// one thread per row, then the records are scaled by sqrt(omega).
template <int U>
__device__ __forceinline__ void stage_tile(const DRole& role, const StageArgs& A, int b, int tile_n, double* rec_val,
                                           uint32_t* rec_idx, int ptid, int pn, double* sc) {
  const int TP = A.TP, T = A.T, inl = A.inline_irls;
  const double* xs = A.xs[b];
  const double* tail = xs + (size_t)A.F * T;
  // explicit shared-space addresses: everything the common kinds touch goes through ld/st.shared
  const uint32_t xs_a = smem_u32(xs), tail_a = smem_u32(tail), rt_a = smem_u32(A.rt), beta_a = smem_u32(A.s_beta);
  const uint32_t lpp_a = smem_u32(A.lp_part), val_a = smem_u32(rec_val), idx_a = smem_u32(rec_idx), recw_a = smem_u32(A.rec_w);
  const int n_items = tile_n * role.n_rterms;
  const float inv_tn = 1.0f / (float)tile_n;
  for (int it0 = ptid; it0 < n_items; it0 += U * pn) {
    double v[U][4];
    int at[U], i0[U], nv[U], gq[U], gr[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int item = it0 + j * pn;
      nv[j] = 0;
      gq[j] = -1;
      if (item >= n_items) continue;
      const int q = (int)(((float)item + 0.5f) * inv_tn);  // item / tile_n, exact for these sizes
      const int rI = item - q * tile_n;
      double lo, inv, dn;
      int kind, nint1, feature, slot0, idx_base, coef_offset;
      ld_stterm(rt_a + q * 64, lo, inv, dn, kind, nint1, feature, slot0, idx_base, coef_offset);
      at[j] = slot0 * TP + rI;
      i0[j] = idx_base;
      double lp = 0.0;
      const double scale = inl ? 1.0 : ld_s64(tail_a + 16 * rI);
      if (kind == kStCubic) {
        const double u = (ld_s64(xs_a + 8 * (feature * T + rI)) - lo) * inv;
        if (u >= 0.0 && u <= 1.0) {
          const double s = u * dn;
          const int i = min((int)s, nint1);
          double b0, b1, b2, b3;
          cubic_bspline(s - (double)i, b0, b1, b2, b3);
          if (inl) {
            const uint32_t bc = beta_a + 8 * (coef_offset + i);
            lp = fma(b3, ld_s64(bc + 24), fma(b2, ld_s64(bc + 16), fma(b1, ld_s64(bc + 8), b0 * ld_s64(bc))));
          }
          v[j][0] = b0 * scale; v[j][1] = b1 * scale; v[j][2] = b2 * scale; v[j][3] = b3 * scale;
          i0[j] += i;
          nv[j] = 4;
        } else {
          gq[j] = q;  // linear extrapolation outside the edge knots: general path
          gr[j] = rI;
        }
      } else if (kind == kStIntercept) {
        if (inl) lp = ld_s64(beta_a + 8 * coef_offset);
        v[j][0] = scale;
        nv[j] = 1;
      } else if (kind == kStLinear) {
        const double x = ld_s64(xs_a + 8 * (feature * T + rI));
        if (inl) lp = x * ld_s64(beta_a + 8 * coef_offset);
        v[j][0] = x * scale;
        nv[j] = 1;
      } else if (kind == kStRhs) {
        if (!inl) { v[j][0] = ld_s64(tail_a + 16 * rI + 8); nv[j] = 1; }
      } else {
        gq[j] = q;
        gr[j] = rI;
      }
      if (inl && kind != kStRhs && gq[j] < 0) st_s64(lpp_a + 8 * (q * T + rI), lp);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (a < nv[j]) {
          st_s64(val_a + 8 * (at[j] + a * TP), v[j][a]);
          st_s32(idx_a + 4 * (at[j] + a * TP), (uint32_t)(i0[j] + a));
        }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (gq[j] < 0) continue;
      const StTerm& t = A.rt[gq[j]];
      const DTerm& d = A.s_terms[t.term];
      const int at2 = t.slot0 * TP + gr[j];
      EmitRec e{rec_val + at2, rec_idx + at2, inl ? A.s_beta : nullptr, 0.0, inl ? 1.0 : tail[2 * gr[j]], TP, d.coef_offset, t.idx_base};
      const LoadTile load{xs, T, gr[j]};
      eval_term(d, load, e);
      if (inl) A.lp_part[gq[j] * T + gr[j]] = e.lp;
    }
  }
  if (inl) {
    bar_sync(kBarProd, pn);
    const int rhs_q = role.n_rterms - 1;  // the rhs is the last role term when present
    for (int rI = ptid; rI < tile_n; rI += pn) {
      double lp = 0.0;
      for (int q = 0; q < role.n_rterms; ++q)
        if (q != rhs_q || role.rhs_slot < 0) lp += ld_s64(lpp_a + 8 * (q * T + rI));
      const double yi = ld_s64(tail_a + 8 * rI);
      double sw, z;
      if (A.mode == PGB_MODE_IRLS) {
        const double winv = A.w ? (double)__fdiv_rn(1.0f, ld_s32f(smem_u32(A.ws[b]) + 4 * rI)) : 1.0;
        const IrlsRow r = irls_row(A.fam, lp, yi, winv);
        sw = r.sw;
        z = r.z;
        if (r.keep) {
          sc[PGB_GS_NUSED] += 1.0;
          sc[PGB_GS_DEV] += dist_dev(A.fam, yi, r.mu);
          sc[PGB_GS_NCORRECT] += (yi == ((r.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
        }
      } else {
        sw = 1.0;
        z = init_response(A.fam, yi);
        sc[PGB_GS_NUSED] += 1.0;
        if (!isfinite(z)) sc[PGB_GS_NONFINITE] += 1.0;
      }
      sc[PGB_GS_NROWS] += 1.0;
      st_s64(recw_a + 8 * rI, sw);
      if (role.rhs_slot >= 0) {
        st_s64(val_a + 8 * (role.rhs_slot * TP + rI), z);
        st_s32(idx_a + 4 * (role.rhs_slot * TP + rI), (uint32_t)A.rt[rhs_q].idx_base);
      }
    }
    bar_sync(kBarProd, pn);
    // scale the records by sqrt(omega); phase B works on row pairs: the odd row out contributes zeros
    const int tn2 = (tile_n + 1) & ~1;
    const float inv_t2 = 1.0f / (float)tn2;
    for (int e = ptid; e < role.n_slots * tn2; e += pn) {
      const int s = (int)(((float)e + 0.5f) * inv_t2), rI = e - s * tn2;
      const uint32_t a = val_a + 8 * (s * TP + rI);
      st_s64(a, (rI < tile_n) ? ld_s64(a) * ld_s64(recw_a + 8 * rI) : 0.0);
    }
  } else if (tile_n & 1) {
    for (int s = ptid; s < role.n_slots; s += pn) st_s64(val_a + 8 * (s * TP + tile_n), 0.0);
  }
}

// Shared-memory record of a tile: for every slot of the role, value = sqrt(omega) * b (the rhs
// slot holds sqrt(omega) * z) and index = position of the coefficient inside its group.
// G = sum (sqrt(omega) b)(sqrt(omega) b)^T is then a plain outer product of the record with
// itself -- the same arithmetic as the reference's W.B (pygam.py:782) before its QR.
// Warps 0 .. role.n_warps-1 are consumers (phase B: accumulation, bound by the shared-memory
// pipe); the remaining warps are producers (phase A: basis, link, weight -> records of the next
// tile, bound by HBM latency and the fp64 pipe).  Two record buffers, named barriers full/empty.
__global__ void __launch_bounds__(32 * (kMaxWarps + kMaxProducers))
gram_accum_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, const DStep* __restrict__ g_steps,
                  const DWarp* __restrict__ g_warps, const DRoleTerm* __restrict__ g_rterms,
                  const DRole* __restrict__ g_roles, int n_roles, Family fam, int mode, int inline_irls,
                  const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta, const double* __restrict__ wz,
                  int T, int F, int dbg, int bulk_ok, double* __restrict__ partials, double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int ridx = 0;
  while (ridx + 1 < n_roles && (int)blockIdx.x >= g_roles[ridx + 1].cta0) ++ridx;
  const DRole role = g_roles[ridx];
  const int cta_in_role = blockIdx.x - role.cta0;
  const int S = role.n_slots, TP = (T + 2) | 1;

  GramSmem sm;
  sm.acc = reinterpret_cast<double*>(smem_raw);
  sm.rec_val[0] = sm.acc + role.acc_size;
  sm.rec_val[1] = sm.rec_val[0] + (size_t)S * TP;
  sm.rec_w = sm.rec_val[1] + (size_t)S * TP;
  sm.lp_part = sm.rec_w + T;  // [n_rterms][T], only when the role computes the IRLS step inline
  sm.s_beta = sm.lp_part + (inline_irls ? (size_t)T * role.n_rterms : 0);
  sm.s_red = sm.s_beta + ((m + 1) & ~1);
  sm.s_terms = reinterpret_cast<DTerm*>(sm.s_red + PGB_GS_COUNT * 32);
  sm.rec_idx[0] = reinterpret_cast<uint32_t*>(sm.s_terms + n_terms);
  sm.rec_idx[1] = sm.rec_idx[0] + (size_t)S * TP;
  sm.steps = reinterpret_cast<DStep*>(sm.rec_idx[0] + (((size_t)2 * S * TP + 1) & ~(size_t)1));
  sm.rt = reinterpret_cast<StTerm*>(sm.steps + role.n_steps);
  sm.xs[0] = reinterpret_cast<double*>(sm.rt + role.n_rterms);
  for (int q = 1; q < kStages; ++q) sm.xs[q] = sm.xs[q - 1] + (size_t)(F + 2) * T;
  sm.ws[0] = reinterpret_cast<float*>(sm.xs[kStages - 1] + (size_t)(F + 2) * T);
  for (int q = 1; q < kStages; ++q) sm.ws[q] = sm.ws[q - 1] + T;
  sm.mbar = reinterpret_cast<unsigned long long*>(sm.ws[kStages - 1] + T);  // T is a multiple of 8: 8-byte aligned

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < role.acc_size; i += blockDim.x) sm.acc[i] = 0.0;
  for (int i = tid; i < 2 * S * TP; i += blockDim.x) { sm.rec_val[0][i] = 0.0; sm.rec_idx[0][i] = 0u; }  // padding slots stay zero
  if (inline_irls) for (int i = tid; i < m; i += blockDim.x) sm.s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(sm.s_terms, g_terms, n_terms);
  if (tid == 0) {
    for (int q = 0; q < kStages; ++q) mbar_init(smem_u32(sm.mbar + q), 1);
    fence_mbar_init();
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_steps + role.step0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm.steps);
    for (int i = tid; i < role.n_steps * (int)(sizeof(DStep) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  for (int i = tid; i < role.n_rterms; i += blockDim.x) {
    const DRoleTerm rt = g_rterms[role.term0 + i];
    StTerm t{};
    t.slot0 = rt.slot0;
    t.idx_base = rt.idx_base;
    t.term = rt.term;
    t.kind = kStRhs;
    if (rt.term >= 0) {
      const DTerm& d = sm.s_terms[rt.term];
      const DMarg& mg = d.marg[0];
      t.kind = d.fast ? kStCubic : d.kind == PGB_TERM_INTERCEPT ? kStIntercept : d.kind == PGB_TERM_LINEAR ? kStLinear : kStGeneric;
      t.lo = mg.lo;
      t.inv = mg.inv_span;
      t.dn = (double)mg.nint;
      t.nint1 = mg.nint - 1;
      t.feature = mg.feature;
      t.coef_offset = d.coef_offset;
    }
    sm.rt[i] = t;
  }
  __syncthreads();

  double sc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) sc[s] = 0.0;

  // rows of this CTA: a contiguous slice of [0, n) in units of T rows
  const int64_t n_tiles = (n + T - 1) / T;
  const int64_t tiles_lo = n_tiles * cta_in_role / role.n_ctas;
  const int64_t tiles_hi = n_tiles * (cta_in_role + 1) / role.n_ctas;
  auto rows_of = [&](int64_t tile) { return (int)min((int64_t)T, n - tile * T); };
  const int nthreads = blockDim.x;

  if (warp >= role.n_warps) {
    // ---------------- producers ----------------
    const int ptid = tid - 32 * role.n_warps, pn = nthreads - 32 * role.n_warps;
    DRole prole = role;
    prole.term0 = 0;  // the role's terms are read from their shared-memory copy
    StageArgs A{sm.rt, g_rterms, sm.s_terms, X, y, w, wz, sm.s_beta, sm.lp_part, sm.rec_w, {}, {}, sm.mbar, ld, n, fam, mode,
                inline_irls, T, TP, F, bulk_ok};
    for (int q = 0; q < kStages; ++q) { A.xs[q] = sm.xs[q]; A.ws[q] = sm.ws[q]; }
    // raw columns are fetched kStages - 1 tiles ahead
    auto fetch = [&](int64_t tile, int b) {
      if (tile < tiles_hi && !(dbg & 4)) stage_fetch(A, tile * T, rows_of(tile), b, ptid, pn);  // dbg 4: no fetch
    };
    for (int q = 0; q < kStages - 1; ++q) fetch(tiles_lo + q, q);
    int buf = 0, sb = 0;
    uint32_t phases = 0u;  // bit b: parity the next wait on stage buffer b expects
    for (int64_t tile = tiles_lo; tile < tiles_hi; ++tile, buf ^= 1, sb = (sb + 1) % kStages) {
      if (!(dbg & 4)) mbar_wait(smem_u32(sm.mbar + sb), (phases >> sb) & 1u);
      phases ^= 1u << sb;
      bar_sync(kBarProd, pn);  // this tile's raw columns have landed; everyone is done with the previous tile's
      fetch(tile + kStages - 1, (sb + kStages - 1) % kStages);
      if (tile >= tiles_lo + 2) bar_sync(kBarEmpty + buf, nthreads);  // the consumers are done with this record buffer
      if (!(dbg & 1) || tile < tiles_lo + 2)  // dbg 1: stage only the first two tiles (consumer-only timing)
        stage_tile<PGB_STAGE_U>(prole, A, sb, rows_of(tile), sm.rec_val[buf], sm.rec_idx[buf], ptid, pn, sc);
      bar_arrive(kBarFull + buf, nthreads);
    }
  } else {
    // ---------------- consumers: every warp walks the tile once per step ----------------
    const DWarp wp = g_warps[role.warp0 + warp];
    const int half = lane >> 4, lu = (lane >> 2) & 3, lt = lane & 3;
    const uint32_t acc_u32 = smem_u32(sm.acc);
    int buf = 0;
    for (int64_t tile = tiles_lo; tile < tiles_hi; ++tile, buf ^= 1) {
      bar_sync(kBarFull + buf, nthreads);
      const int tile_n = rows_of(tile);
      const uint32_t val_u32 = smem_u32(sm.rec_val[buf]), idx_u32 = smem_u32(sm.rec_idx[buf]);
      for (int s = 0; s < ((dbg & 2) ? 0 : wp.n_steps); ++s) {  // dbg 2: no accumulation (producer-only timing)
        const DHalf* hp = &sm.steps[wp.step0 - role.step0 + s].h[half];
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
      if (tile + 2 < tiles_hi) bar_arrive(kBarEmpty + buf, nthreads);
    }
  }
  __syncthreads();
  // ---------------- write the partial ----------------
  double* dst = partials + role.part_off + (int64_t)cta_in_role * role.acc_size;
  for (int64_t i = tid; i < role.acc_size; i += blockDim.x) dst[i] = sm.acc[i];
  if (inline_irls) block_reduce_store<PGB_GS_COUNT>(sc, sm.s_red, scal_part + (int64_t)blockIdx.x * PGB_GS_COUNT);
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
  if (!M || !X || !y || !out || !ws || n < 1 || ld < n || (mode == PGB_MODE_IRLS && !beta) ||
      (mode != PGB_MODE_IRLS && mode != PGB_MODE_INIT)) {
    pgb_set_error("pgb_gram_pass: bad arguments");
    return PGB_EINVAL;
  }
  if (ws_bytes < pgb_gram_ws_bytes(M, n)) {
    pgb_set_error("pgb_gram_pass: workspace too small (%lld < %lld)", (long long)ws_bytes,
                  (long long)pgb_gram_ws_bytes(M, n));
    return PGB_EINVAL;
  }
  const int n_roles = (int)M->roles.size();
  double* partials = (double*)ws;
  double* scal_part = partials + M->partial_doubles;
  double* wz = scal_part + (int64_t)(M->gram_ctas + M->pass_ctas) * PGB_GS_COUNT;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PGB_CUDA(cudaFuncSetAttribute(gram_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    PGB_CUDA(cudaFuncSetAttribute(irls_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    attr_set = true;
  }
  const int inline_irls = gram_inline(M) ? 1 : 0;
  int scal_count = M->gram_ctas;
  if (!inline_irls) {
    const int grid = (int)std::min<int64_t>((n + kPassThreads - 1) / kPassThreads, M->pass_ctas);
    const size_t smem = (size_t)((M->m + 1) & ~1) * 8 + (size_t)M->n_terms * sizeof(DTerm) + PGB_GS_COUNT * 32 * 8;
    irls_rows_kernel<<<grid, kPassThreads, smem, st>>>(M->d_terms, M->n_terms, M->m, M->fam, mode, X, n, ld, y,
                                                       w, beta, wz, scal_part);
    pgb_count_launch(1);
    PGB_CUDA(cudaGetLastError());
    scal_count = grid;
  }
  const int bulk_ok = ((((uintptr_t)X | (uintptr_t)y | (uintptr_t)w | (uintptr_t)wz) & 15) == 0 && (ld & 1) == 0) ? 1 : 0;  // TMA bulk copies need 16-byte alignment
  const char* edbg = getenv("PGB_GRAM_DEBUG");  // timing experiments only
  const int dbg = edbg ? atoi(edbg) : 0;
  pgb_model::Timing& tm = const_cast<pgb_model*>(M)->timing;
  if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.ev[0], st));
  gram_accum_kernel<<<M->gram_ctas, M->gram_threads, M->gram_smem, st>>>(
      M->d_terms, M->n_terms, M->m, M->d_steps, M->d_warps, M->d_rterms, M->d_roles, n_roles, M->fam, mode,
      inline_irls, X, n, ld, y, w, beta, wz, M->tile_rows, M->n_features, dbg, bulk_ok, partials, scal_part);
  if (tm.enabled) { PGB_CUDA(cudaEventRecord(tm.ev[1], st)); tm.pending = true; }
  PGB_CUDA(cudaGetLastError());
  const int rb = 256;
  gram_reduce_kernel<<<(unsigned)((M->map_size + rb - 1) / rb), rb, 0, st>>>(
      partials, M->d_roles, n_roles, M->d_elem_map, M->map_size, M->m, out, accumulate);
  scalar_reduce_kernel<<<1, 32, 0, st>>>(scal_part, scal_count, PGB_GS_COUNT,
                                         out + (int64_t)M->m * M->m + M->m, accumulate);
  pgb_count_launch(3);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// optional timing of the dominant kernel (gram_accum_kernel) with CUDA events recorded on the
// launching stream; used by bench.py for the roofline figure
extern "C" int pgb_gram_timing(pgb_model* M, int32_t enable) {
  if (!M) { pgb_set_error("pgb_gram_timing: null model"); return PGB_EINVAL; }
  pgb_model::Timing& tm = M->timing;
  if (enable && !tm.ev[0]) {
    PGB_CUDA(cudaEventCreate(&tm.ev[0]));
    PGB_CUDA(cudaEventCreate(&tm.ev[1]));
  }
  tm.enabled = enable != 0;
  tm.pending = false;
  return PGB_OK;
}
extern "C" int pgb_gram_last_ms(pgb_model* M, float* accum_ms) {
  if (!M || !accum_ms || !M->timing.pending) { pgb_set_error("pgb_gram_last_ms: no timed pass"); return PGB_EINVAL; }
  PGB_CUDA(cudaEventSynchronize(M->timing.ev[1]));
  PGB_CUDA(cudaEventElapsedTime(accum_ms, M->timing.ev[0], M->timing.ev[1]));
  return PGB_OK;
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
  S.ws_bytes = pgb_gram_ws_bytes(M, chunk_rows);
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
