// pgb_gram.cu -- the fused Gram pass of one PIRLS iteration.
//
// Replaces pygam.py:764-783 of the reference (lp -> mu -> W -> mask -> pseudo data -> W.B and
// the n-sized part of its QR) and the data part of _initial_estimate (pygam.py:696-710):
//     G = X^T W^2 X,   r = X^T W^2 z,   plus the per-iteration log scalars,
// with the model-matrix rows evaluated on the fly (never materialised).
//
// Work decomposition (pgb_internal.h): a UNIT is the block G[coefficients of term j, columns of
// chunk g]; one warp owns one unit for the whole kernel and keeps it in shared memory, so the
// fp64 read-modify-write of the accumulators needs no atomics.  A CTA (role) stages a tile of
// rows as (position, value) records in shared memory (phase A, one thread per row: basis,
// link, weight), then every warp walks the tile and adds the row's weighted outer product into
// its unit (phase B): lanes = 2 row-side non-zeros x 16 column-side slots.  With the 4-way
// interleaved column layout of cubic-spline chunks the 16 lanes of a half-warp hit 16 distinct
// 8-byte banks.  Each CTA writes its accumulators to a private partial; gram_reduce_kernel sums
// the partials in a fixed order (bit-reproducible) and mirrors the upper triangle.
#include <cuda_runtime.h>
#include <stdio.h>

#include <algorithm>

#include "pgb_internal.h"
#include "pgb_kernels.cuh"

static const int kMaxUnitsPerRole = 16;  // warps per CTA
static const int kPlanTile = 64;         // tile rows assumed while packing roles

// ----------------------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------------------
static bool term_interleavable(const DTerm& t) {
  if (t.kind == PGB_TERM_TENSOR) return false;
  if (t.nnz > 4) return false;
  if (t.kind == PGB_TERM_SPLINE && t.marg[0].periodic) return false;  // wrapped columns
  return true;
}

static size_t role_fixed_smem(const pgb_model* M) {
  return (size_t)M->n_terms * sizeof(DTerm) + (size_t)M->m * 8 + PGB_GS_COUNT * 32 * 8 + 512;
}
static size_t rec_row_bytes(int n_slots) { return (size_t)n_slots * 12 + 8; }
static int padded_nnz(const DTerm& t) { return (t.nnz + 1) & ~1; }  // record slots of a term (even)

int pgb_plan_gram(pgb_model* M) {
  const int m = M->m, nt = M->n_terms;
  std::vector<DTerm>& T = M->dterms;
  // ---- column chunks: consecutive terms, at most 16 slots, at most 4 terms when interleaved
  M->chunks.clear();
  int t = 0;
  while (t < nt) {
    DChunk c{};
    c.first_term = t;
    const bool inter = term_interleavable(T[t]);
    int slots = padded_nnz(T[t]), nterms = 1;
    ++t;
    while (t < nt && slots + padded_nnz(T[t]) <= 16) {
      if (inter && (!term_interleavable(T[t]) || nterms >= 4)) break;
      if (!inter && nterms >= 16) break;
      slots += padded_nnz(T[t]);
      ++nterms;
      ++t;
    }
    c.n_terms = nterms;
    c.n_slots = slots;
    c.layout = inter ? 1 : 0;
    M->chunks.push_back(c);
  }
  // the virtual rhs column joins the last chunk when a slot is free
  if (M->chunks.back().n_slots + 1 > 16) {
    DChunk c{};
    c.first_term = nt;
    c.n_terms = 0;
    c.layout = 0;
    M->chunks.push_back(c);
  }
  M->chunks.back().has_rhs = 1;
  M->chunks.back().n_slots += 1;
  for (size_t g = 0; g < M->chunks.size(); ++g) {
    DChunk& c = M->chunks[g];
    int width = 0, maxk = 0;
    for (int q = 0; q < c.n_terms; ++q) {
      DTerm& d = T[c.first_term + q];
      d.chunk = (int)g;
      d.pos_mode = c.layout;
      d.pos_base = c.layout ? 4 * q : width;
      width += d.n_coefs;
      maxk = std::max(maxk, d.n_coefs);
    }
    if (c.layout) width = 16 * ((maxk + 3) / 4);
    c.pad0 = width;  // position of the rhs column: appended after the chunk's own columns
    c.width = width + (c.has_rhs ? 1 : 0);
  }
  // ---- units in (row term, chunk) order
  std::vector<DUnit> all;
  for (int j = 0; j < nt; ++j) {
    for (int g = T[j].chunk; g < (int)M->chunks.size(); ++g) {
      const DChunk& c = M->chunks[g];
      DUnit u{};
      u.row_term = j;
      u.chunk = g;
      u.nnz_r = T[j].nnz;
      u.clen = c.n_slots;
      u.cmin = 0;
      if (g == T[j].chunk)
        for (int q = c.first_term; q < j; ++q) u.cmin += padded_nnz(T[q]);
      u.width = c.width;
      u.k_rows = T[j].n_coefs;
      all.push_back(u);
    }
  }
  // ---- pack units into roles
  const size_t fixed = role_fixed_smem(M);
  int k_pad = 6;
  for (const DTerm& d : T) k_pad += padded_nnz(d);
  const size_t rec_max = rec_row_bytes(k_pad);
  if (fixed + (size_t)32 * rec_max + 8192 > kSmemMax) {
    pgb_set_error("model too wide for the Gram pass: k=%d non-zeros per row", M->k);
    return PGB_EUNSUPPORTED;
  }
  int plan_tile = kPlanTile;
  while (plan_tile > 32 && fixed + (size_t)plan_tile * rec_max > kSmemMax / 2) plan_tile >>= 1;
  const size_t acc_budget = (kSmemMax - fixed - (size_t)(plan_tile + 1) * rec_max) / 8;  // doubles
  M->units.clear();
  M->roles.clear();
  M->rterms.clear();
  size_t i = 0;
  while (i < all.size()) {
    DRole r{};
    r.unit0 = (int)M->units.size();
    int64_t acc = 0;
    while (i < all.size() && r.n_units < kMaxUnitsPerRole) {
      const int64_t sz = (int64_t)all[i].k_rows * all[i].width;
      if ((size_t)sz > acc_budget) {
        pgb_set_error("a Gram block of %d x %d doubles does not fit in shared memory", all[i].k_rows, all[i].width);
        return PGB_EUNSUPPORTED;
      }
      if (r.n_units > 0 && (size_t)(acc + sz) > acc_budget) break;
      DUnit u = all[i];
      u.off = acc;
      acc += sz;
      M->units.push_back(u);
      ++r.n_units;
      ++i;
    }
    r.acc_size = acc;
    // terms the role must evaluate: its row terms and every term of its chunks
    std::vector<char> need(nt, 0);
    bool need_rhs = false;
    for (int q = 0; q < r.n_units; ++q) {
      const DUnit& u = M->units[r.unit0 + q];
      need[u.row_term] = 1;
      const DChunk& c = M->chunks[u.chunk];
      for (int w = 0; w < c.n_terms; ++w) need[c.first_term + w] = 1;
      if (c.has_rhs) need_rhs = true;
    }
    r.term0 = (int)M->rterms.size();
    int slot = 0;
    std::vector<int> slot_of(nt, -1);
    bool all_terms = true;
    for (int q = 0; q < nt; ++q) {
      if (!need[q]) { all_terms = false; continue; }
      M->rterms.push_back(DRoleTerm{q, slot});
      slot_of[q] = slot;
      slot += padded_nnz(T[q]);
    }
    int rhs_slot = -1;
    if (need_rhs) {
      M->rterms.push_back(DRoleTerm{-1, slot});
      rhs_slot = slot;
      slot += 2;
    }
    r.n_rterms = (int)M->rterms.size() - r.term0;
    while ((slot & 3) != 2) slot += 2;  // even (16-byte aligned rows) and 2 mod 4 (row-strided stores spread over banks)
    r.n_slots = slot;
    r.needs_all_terms = all_terms ? 1 : 0;
    r.rhs_pos = M->chunks.back().pad0;
    r.rhs_slot = rhs_slot;
    for (int q = 0; q < r.n_units; ++q) {
      DUnit& u = M->units[r.unit0 + q];
      u.rslot = slot_of[u.row_term];
      const DChunk& c = M->chunks[u.chunk];
      u.cslot = c.n_terms > 0 ? slot_of[c.first_term] : rhs_slot;
    }
    M->roles.push_back(r);
  }
  // ---- tile rows, shared memory, threads: the largest tile that fits; if a tile of >= 96 rows
  //      also fits twice per SM, take that one (two resident CTAs hide each other's phase A)
  int max_units = 1;
  for (const DRole& r : M->roles) max_units = std::max(max_units, r.n_units);
  auto smem_for = [&](int tile) {
    size_t smem = 0;
    for (const DRole& r : M->roles) {
      size_t need = (size_t)(r.acc_size + 1) * 8 + (size_t)(tile + 1) * rec_row_bytes(r.n_slots) + fixed;
      if (M->roles.size() == 1) need += (size_t)tile * r.n_rterms * 8;  // lp_part of the inline IRLS step
      smem = std::max(smem, need);
    }
    return smem;
  };
  int tile = 256;
  while (tile > 32 && smem_for(tile) > kSmemMax) tile -= 32;
  for (int t2 = tile; t2 >= 96; t2 -= 32) {
    if (smem_for(t2) + 1024 <= kSmemMax / 2) { tile = t2; break; }
  }
  size_t smem = smem_for(tile);
  if (smem > kSmemMax) {
    pgb_set_error("Gram pass does not fit in shared memory (%zu bytes)", smem);
    return PGB_EUNSUPPORTED;
  }
  M->tile_rows = tile;
  M->gram_smem = (smem + 127) & ~(size_t)127;
  M->gram_threads = std::max(128, 32 * max_units);
  // ---- CTAs per role proportional to the role's phase-B cost
  int occ = (int)std::min<size_t>(8, kSmemMax / (M->gram_smem + 1024));
  occ = std::max(1, std::min(occ, 2048 / M->gram_threads));
  const int n_roles = (int)M->roles.size();
  const int total = std::max(kSMs * occ, n_roles);
  std::vector<double> cost(n_roles, 0.0);
  double cost_sum = 0.0;
  for (int q = 0; q < n_roles; ++q) {
    const DRole& r = M->roles[q];
    double mx = 0.0;
    for (int w = 0; w < r.n_units; ++w) {
      const DUnit& u = M->units[r.unit0 + w];
      mx = std::max(mx, (double)((u.nnz_r + 1) / 2) * ((u.clen + 15) / 16));
    }
    cost[q] = mx * 13.0 + 0.15 * r.n_slots * 12.0 / std::max(1, M->gram_threads / 32) + 4.0;
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
  // ---- element map: accumulator element -> row * (m + 1) + col of [G | r], upper triangle only
  M->elem_map.assign((size_t)moff, -1);
  for (const DRole& r : M->roles) {
    for (int q = 0; q < r.n_units; ++q) {
      const DUnit& u = M->units[r.unit0 + q];
      const DChunk& c = M->chunks[u.chunk];
      const DTerm& rt = T[u.row_term];
      for (int ri = 0; ri < u.k_rows; ++ri) {
        const int grow = rt.coef_offset + ri;
        int32_t* dst = M->elem_map.data() + r.map_off + u.off + (int64_t)ri * u.width;
        for (int w = 0; w < c.n_terms; ++w) {
          const int ct = c.first_term + w;
          if (ct < u.row_term) continue;
          const DTerm& d = T[ct];
          for (int ci = 0; ci < d.n_coefs; ++ci) {
            const int gcol = d.coef_offset + ci;
            if (gcol < grow) continue;
            dst[term_pos(d.pos_base, d.pos_mode, ci)] = grow * (m + 1) + gcol;
          }
        }
        if (c.has_rhs) dst[c.pad0] = grow * (m + 1) + m;
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
  up((void**)&M->d_chunks, M->chunks.data(), M->chunks.size() * sizeof(DChunk));
  up((void**)&M->d_units, M->units.data(), M->units.size() * sizeof(DUnit));
  up((void**)&M->d_rterms, M->rterms.data(), M->rterms.size() * sizeof(DRoleTerm));
  up((void**)&M->d_roles, M->roles.data(), M->roles.size() * sizeof(DRole));
  up((void**)&M->d_elem_map, M->elem_map.data(), M->elem_map.size() * sizeof(int32_t));
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
  return errors;
}

static void stage_free(pgb_model* M);

void pgb_gram_release(pgb_model* M) {
  stage_free(M);
  if (M->timing.ev[0]) { cudaEventDestroy(M->timing.ev[0]); cudaEventDestroy(M->timing.ev[1]); }
  cudaFree(M->d_chunks);
  cudaFree(M->d_units);
  cudaFree(M->d_rterms);
  cudaFree(M->d_roles);
  cudaFree(M->d_elem_map);
  M->d_chunks = nullptr; M->d_units = nullptr; M->d_rterms = nullptr; M->d_roles = nullptr; M->d_elem_map = nullptr;
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
__device__ __forceinline__ void lds_v2f64(double& x, double& y, uint32_t a) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds_v2u32(uint32_t& x, uint32_t& y, uint32_t a) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// predicated forms: a lane whose predicate is off issues no shared-memory access at all
__device__ __forceinline__ void lds_f64_if(double& v, uint32_t a, uint32_t on) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.shared.f64 %0, [%1]; }" : "+d"(v) : "r"(a), "r"(on));
}
__device__ __forceinline__ void sts_f64_if(uint32_t a, double v, uint32_t on) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"(on)
               : "memory");
}

struct EmitRec {
  double* val;
  uint32_t* pr;
  const double* beta;
  double lp, scale;
  int q, coef_offset, pos_base, pos_mode;
  __device__ __forceinline__ void operator()(int c, double v) {
    const int i = c - coef_offset;
    val[q] = v * scale;
    pr[q] = ((uint32_t)term_pos(pos_base, pos_mode, i) << 3) | ((uint32_t)i << 16);  // 8 * column | row
    ++q;
    if (beta) lp = fma(v, beta[c], lp);
  }
};

// Shared-memory row record of a tile: for every slot of the role, value[slot] = sqrt(omega) * b
// (the rhs slot holds sqrt(omega) * z) and pr[slot] = physical column | coefficient index << 16.
// G = sum (sqrt(omega) b)(sqrt(omega) b)^T is then a plain outer product of the record with
// itself -- the same arithmetic as the reference's W.B (pygam.py:782) before its QR.
__global__ void __launch_bounds__(512)
gram_accum_kernel(const DTerm* __restrict__ g_terms, int n_terms, int m, const DChunk* __restrict__ g_chunks,
                  const DUnit* __restrict__ g_units, const DRoleTerm* __restrict__ g_rterms,
                  const DRole* __restrict__ g_roles, int n_roles, Family fam, int mode, int inline_irls,
                  const double* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ y,
                  const float* __restrict__ w, const double* __restrict__ beta, const double* __restrict__ wz,
                  int T, double* __restrict__ partials, double* __restrict__ scal_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int ridx = 0;
  while (ridx + 1 < n_roles && (int)blockIdx.x >= g_roles[ridx + 1].cta0) ++ridx;
  const DRole role = g_roles[ridx];
  const int cta_in_role = blockIdx.x - role.cta0;
  const int S = role.n_slots;  // even

  double* acc = reinterpret_cast<double*>(smem_raw);
  double* rec_val = acc + ((role.acc_size + 1) & ~(int64_t)1);  // 16-byte aligned
  double* rec_w = rec_val + (size_t)(T + 1) * S;  // one spare row: the prefetch of phase B runs one ahead
  double* lp_part = rec_w + T;  // [n_rterms][T], only when the role computes the IRLS step inline
  double* s_beta = lp_part + (inline_irls ? (size_t)T * role.n_rterms : 0);
  double* s_red = s_beta + ((m + 1) & ~1);
  DTerm* s_terms = reinterpret_cast<DTerm*>(s_red + PGB_GS_COUNT * 32);
  uint32_t* rec_pr = reinterpret_cast<uint32_t*>(s_terms + n_terms);  // (T + 1) rows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < role.acc_size; i += blockDim.x) acc[i] = 0.0;
  for (int i = tid; i < (T + 1) * S; i += blockDim.x) { rec_val[i] = 0.0; rec_pr[i] = 0u; }  // pad slots stay zero
  if (inline_irls) for (int i = tid; i < m; i += blockDim.x) s_beta[i] = (mode == PGB_MODE_IRLS) ? beta[i] : 0.0;
  load_terms_smem(s_terms, g_terms, n_terms);
  __syncthreads();

  double sc[PGB_GS_COUNT];
#pragma unroll
  for (int s = 0; s < PGB_GS_COUNT; ++s) sc[s] = 0.0;

  // rows of this CTA: a contiguous slice of [0, n) in units of T rows
  const int64_t n_tiles = (n + T - 1) / T;
  const int64_t tiles_lo = n_tiles * cta_in_role / role.n_ctas;
  const int64_t tiles_hi = n_tiles * (cta_in_role + 1) / role.n_ctas;

  const bool has_unit = warp < role.n_units;
  DUnit u{};
  if (has_unit) u = g_units[role.unit0 + warp];
  const int s16 = lane & 15, h = lane >> 4;
  double* uacc = acc + u.off;
  const int width = u.width;
  const bool fast = has_unit && u.clen <= 16 && u.nnz_r <= 4;
  const bool cvalid = has_unit && (s16 < u.clen) && (s16 >= u.cmin);
  const bool v0 = cvalid && (2 * h < u.nnz_r), v1 = cvalid && (2 * h + 1 < u.nnz_r);
  const int coff = u.cslot + s16, roff = u.rslot + 2 * h;
  const uint32_t smem_cv = smem_u32(rec_val + coff), smem_cp = smem_u32(rec_pr + coff);
  const uint32_t smem_rv = smem_u32(rec_val + roff), smem_rp = smem_u32(rec_pr + roff);
  const uint32_t acc_u32 = smem_u32(uacc), w8 = (uint32_t)width * 8u, sb8 = (uint32_t)S * 8u, sb4 = (uint32_t)S * 4u;
  const uint32_t p0 = v0 ? 1u : 0u, p1 = v1 ? 1u : 0u;

  for (int64_t tile = tiles_lo; tile < tiles_hi; ++tile) {
    const int64_t row0 = tile * T;
    const int tile_n = (int)min((int64_t)T, n - row0);
    // ---------------- phase A: one thread per (row, term) ----------------
    // consecutive threads take consecutive rows of the same term: coalesced feature loads and
    // every warp of the CTA busy even when the tile is smaller than the block
    const int n_items = tile_n * role.n_rterms;
    for (int item = tid; item < n_items; item += blockDim.x) {
      const int q = item / tile_n;
      const int rI = item - q * tile_n;
      const DRoleTerm rt = g_rterms[role.term0 + q];
      const int64_t i = row0 + rI;
      if (rt.term < 0) {  // rhs slot: sqrt(omega) * z (split mode; inline mode fills it below)
        if (!inline_irls) {
          rec_val[(size_t)rI * S + rt.slot0] = __ldg(wz + 2 * i + 1);
          rec_pr[(size_t)rI * S + rt.slot0] = (uint32_t)role.rhs_pos << 3;
        }
        continue;
      }
      const DTerm& d = s_terms[rt.term];
      EmitRec e{rec_val + (size_t)rI * S, rec_pr + (size_t)rI * S, inline_irls ? s_beta : nullptr, 0.0,
                inline_irls ? 1.0 : __ldg(wz + 2 * i), rt.slot0, d.coef_offset, d.pos_base, d.pos_mode};
      eval_term(d, X, ld, i, e);
      if (inline_irls) lp_part[(size_t)q * T + rI] = e.lp;
    }
    if (inline_irls) {
      __syncthreads();
      for (int rI = tid; rI < tile_n; rI += blockDim.x) {
        const int64_t i = row0 + rI;
        double lp = 0.0;
        for (int q = 0; q < role.n_rterms; ++q)
          if (g_rterms[role.term0 + q].term >= 0) lp += lp_part[(size_t)q * T + rI];
        const double yi = __ldg(y + i);
        double sw, z;
        if (mode == PGB_MODE_IRLS) {
          const IrlsRow r = irls_row(fam, lp, yi, weight_inv(w, i));
          sw = r.sw;
          z = r.z;
          if (r.keep) {
            sc[PGB_GS_NUSED] += 1.0;
            sc[PGB_GS_DEV] += dist_dev(fam, yi, r.mu);
            sc[PGB_GS_NCORRECT] += (yi == ((r.mu > 0.5) ? 1.0 : 0.0)) ? 1.0 : 0.0;
          }
        } else {
          sw = 1.0;
          z = init_response(fam, yi);
          sc[PGB_GS_NUSED] += 1.0;
          if (!isfinite(z)) sc[PGB_GS_NONFINITE] += 1.0;
        }
        sc[PGB_GS_NROWS] += 1.0;
        rec_w[rI] = sw;
        if (role.rhs_slot >= 0) {
          rec_val[(size_t)rI * S + role.rhs_slot] = z;
          rec_pr[(size_t)rI * S + role.rhs_slot] = (uint32_t)role.rhs_pos << 3;
        }
      }
      __syncthreads();
      // scale the row's record by sqrt(omega)
      for (int e = tid; e < tile_n * S; e += blockDim.x) rec_val[e] *= rec_w[e / S];
    }
    __syncthreads();
    // ---------------- phase B: one warp per unit ----------------
    if (fast) {
      // lanes = 16 column slots x 2 row-side pairs (a = 2h, 2h + 1).  Explicit shared-space
      // loads / stores on 32-bit addresses; lanes without work are predicated off (no branch, no
      // shared-memory traffic); the operands of the next row are fetched while the
      // read-modify-write of the current row is in flight.
      asm volatile("" ::: "memory");
      uint32_t a_cv = smem_cv, a_cp = smem_cp, a_rv = smem_rv, a_rp = smem_rp;
      double cv, rv0, rv1;
      uint32_t cpr, rp0, rp1;
      lds_f64(cv, a_cv);
      lds_u32(cpr, a_cp);
      lds_v2f64(rv0, rv1, a_rv);
      lds_v2u32(rp0, rp1, a_rp);
      for (int rI = 0; rI < tile_n; ++rI) {
        const uint32_t bp = acc_u32 + (cpr & 0xffffu);  // accumulator base + 8 * column
        const uint32_t g0 = (rp0 >> 16) * w8 + bp, g1 = (rp1 >> 16) * w8 + bp;
        const double c_cv = cv, c_rv0 = rv0, c_rv1 = rv1;
        a_cv += sb8; a_cp += sb4; a_rv += sb8; a_rp += sb4;  // the record buffer has one spare row
        lds_f64(cv, a_cv);
        lds_u32(cpr, a_cp);
        lds_v2f64(rv0, rv1, a_rv);
        lds_v2u32(rp0, rp1, a_rp);
        double x0 = 0.0, x1 = 0.0;
        lds_f64_if(x0, g0, p0);
        lds_f64_if(x1, g1, p1);
        x0 = fma(c_rv0, c_cv, x0);
        x1 = fma(c_rv1, c_cv, x1);
        sts_f64_if(g0, x0, p0);
        sts_f64_if(g1, x1, p1);
      }
    } else if (has_unit) {
      const int ncp = (u.clen + 15) >> 4;
      for (int rI = 0; rI < tile_n; ++rI) {
        const double* rv = rec_val + (size_t)rI * S;
        const uint32_t* rp = rec_pr + (size_t)rI * S;
        for (int cp = 0; cp < ncp; ++cp) {
          const int cs = (cp << 4) + s16;
          const bool valid = (cs < u.clen) && (cs >= u.cmin);
          const int pos = valid ? (int)((rp[u.cslot + cs] & 0xffffu) >> 3) : 0;
          const double vv = valid ? rv[u.cslot + cs] : 0.0;
          for (int a = h; a < u.nnz_r; a += 2) {
            const int ri = (int)(rp[u.rslot + a] >> 16);
            const double wq = rv[u.rslot + a];
            if (valid) {
              double* g = uacc + ri * width + pos;
              *g = fma(wq, vv, *g);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---------------- write the partial ----------------
  double* dst = partials + role.part_off + (int64_t)cta_in_role * role.acc_size;
  for (int64_t i = tid; i < role.acc_size; i += blockDim.x) dst[i] = acc[i];
  if (inline_irls) block_reduce_store<PGB_GS_COUNT>(sc, s_red, scal_part + (int64_t)blockIdx.x * PGB_GS_COUNT);
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
  pgb_model::Timing& tm = const_cast<pgb_model*>(M)->timing;
  if (tm.enabled) PGB_CUDA(cudaEventRecord(tm.ev[0], st));
  gram_accum_kernel<<<M->gram_ctas, M->gram_threads, M->gram_smem, st>>>(
      M->d_terms, M->n_terms, M->m, M->d_chunks, M->d_units, M->d_rterms, M->d_roles, n_roles, M->fam, mode,
      inline_irls, X, n, ld, y, w, beta, wz, M->tile_rows, partials, scal_part);
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
